"""The drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/wealy_b200.h declares, and the Python mirror fails loudly instead of falling back to CPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "wealy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wealy_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import wealy_b200._native as N
    names = _declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(N.lib, n), f"{n} declared in include/wealy_b200.h but not exported"
        assert n in N.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(N.SIGNATURES) == names


def test_version_and_workspace_queries_need_no_gpu():
    import wealy_b200._native as N
    assert N.lib.wealy_version() >= 100
    assert N.lib.wealy_sim_matrix_workspace_bytes(1000, 2000, 1024, 3) >= (1000 + 2000) * 1024 * 4
    assert N.lib.wealy_sim_matrix_workspace_bytes(1000, 2000, 1024, 1) < N.lib.wealy_sim_matrix_workspace_bytes(1000, 2000, 1024, 3)
    assert N.lib.wealy_loss_workspace_bytes(4096, 1024, 3) > 4096 * 4096 * 4
    assert N.lib.wealy_loss_workspace_bytes(0, 1024, 3) == 0


def test_native_library_is_the_in_tree_build():
    import wealy_b200._native as N
    assert os.path.samefile(os.path.dirname(N._LIB_PATH), os.path.join(ROOT, "audio-based-lyrics-matching_b200", "lib"))
    sass_markers = open(N._LIB_PATH, "rb").read()
    assert b"sm_100a" in sass_markers


def test_no_cpu_fallback():
    from wealy_b200 import tensor_ops as wt, losses as wl
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wt.pairwise_distance_matrix(x, x, mode="cossim")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wl.NTXentLoss()(torch.arange(4), torch.arange(4), x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wl.CLEWSLoss()(torch.arange(4), torch.arange(4), x)


def test_reference_error_behaviour_before_any_compute():
    from wealy_b200 import tensor_ops as wt, losses as wl
    with pytest.raises(AssertionError):                      # lib/tensor_ops.py:153
        wt.pairwise_distance_matrix(torch.zeros(2, 2, 2), torch.zeros(2, 2, 2))
    with pytest.raises(AssertionError):
        wt.pairwise_distance_matrix(torch.zeros(2, 2), torch.zeros(2))
    with pytest.raises(AssertionError):                      # lib/losses.py:31
        wl.NTXentLoss()(torch.arange(3), torch.arange(4), torch.zeros(4, 2))
    with pytest.raises(AssertionError):                      # lib/losses.py:218 (B >= 4)
        wl.CLEWSLoss()(torch.arange(3), torch.arange(3), torch.zeros(3, 2))
    with pytest.raises(AssertionError):                      # lib/losses.py:214 (S must be 1)
        wl.CLEWSLoss()(torch.arange(4), torch.arange(4), torch.zeros(4, 2, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "audio-based-lyrics-matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                if f != "make_clique_sizes.py":     # docstrings may CITE the reference, nothing may read it
                    assert not re.search(r"(open|listdir|insert|exists|isfile)\([^)]*reference", src), f


def test_ctypes_signatures_match_the_header_argument_counts():
    """Every prototype in include/wealy_b200.h has as many parameters as its ctypes signature in _native.py."""
    import wealy_b200._native as N
    src = open(os.path.join(ROOT, "include", "wealy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\b(wealy_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(protos) == set(N.SIGNATURES)
    for name, params in protos.items():
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(N.SIGNATURES[name][1]), f"{name}: header has {n} parameters, ctypes {len(N.SIGNATURES[name][1])}"
