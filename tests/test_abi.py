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


def _ctype_of(decl):
    """C parameter / return declaration from include/wealy_b200.h -> the ctypes type a binding must use."""
    import ctypes
    decl = " ".join(decl.replace("const", " ").split())
    name_stripped = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", decl).strip() if not decl.endswith("*") else decl
    t = name_stripped if name_stripped else decl          # (a bare type without a parameter name)
    t = t.replace(" *", "*").strip()
    stars = t.count("*")
    base = t.replace("*", "").strip()
    scalars = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "float": ctypes.c_float,
               "double": ctypes.c_double, "size_t": ctypes.c_size_t, "uint32_t": ctypes.c_uint32}
    if stars == 0:
        return scalars[base]
    if base == "char" and stars == 1:
        return ctypes.c_char_p
    return "pointer:" + base + "*" * stars


def _compatible(c_decl_type, ctypes_type):
    """A pointer parameter may be bound as c_void_p (raw device / handle pointers) or as POINTER(T) of its pointee
    (host out-parameters, the config struct); a scalar must be bound with exactly its C type."""
    import ctypes
    if not isinstance(c_decl_type, str):
        return c_decl_type is ctypes_type
    if ctypes_type is ctypes.c_void_p:
        return True
    if not hasattr(ctypes_type, "_type_"):
        return False
    base, stars = c_decl_type[len("pointer:"):].rstrip("*"), c_decl_type.count("*")
    pointee = ctypes_type._type_
    if stars >= 2:                                           # T** out-parameter: POINTER(c_void_p)
        return pointee is ctypes.c_void_p
    table = {"int64_t": ctypes.c_int64, "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double,
             "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32}
    if base in table:
        return pointee is table[base]
    return base in ("wealy_loss_cfg", "wealy_eval_plan") and issubclass(pointee, ctypes.Structure) or pointee is ctypes.c_void_p


def test_ctypes_signatures_match_the_header_types():
    """Every prototype in include/wealy_b200.h against its ctypes signature in _native.py: same symbol set, same
    parameter count, and parameter by parameter the same C type (a swapped int64_t / int pair must not pass)."""
    import ctypes
    import wealy_b200._native as N
    src = open(os.path.join(ROOT, "include", "wealy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)\b(wealy_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    names = [n for _, n, _ in protos]
    assert set(names) == set(N.SIGNATURES) and len(names) == len(set(names))
    checked = 0
    for ret, name, params in protos:
        res, args = N.SIGNATURES[name]
        ret = " ".join(ret.split())
        if ret == "void":
            assert res is None, name
        else:
            assert _compatible(_ctype_of(ret + " x") if not ret.endswith("*") else _ctype_of(ret), res) or \
                (ret.replace("const", "").strip() == "char*" and res is ctypes.c_char_p), (name, ret, res)
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, ctypes {len(args)}"
        for k, (decl, ct) in enumerate(zip(plist, args)):
            assert _compatible(_ctype_of(decl), ct), f"{name}: parameter {k} is `{decl}` in the header but {ct} in _native.py"
            checked += 1
    assert checked > 300


def test_type_check_catches_a_swapped_pair():
    import ctypes
    assert not _compatible(_ctype_of("int64_t n"), ctypes.c_int)
    assert not _compatible(_ctype_of("int mode"), ctypes.c_int64)
    assert not _compatible(_ctype_of("float eps"), ctypes.c_double)
    assert not _compatible(_ctype_of("const void* x"), ctypes.c_int64)
    assert _compatible(_ctype_of("const int64_t* offsets"), ctypes.c_void_p)
    assert _compatible(_ctype_of("int64_t* total_pairs"), ctypes.POINTER(ctypes.c_int64))
    assert not _compatible(_ctype_of("int64_t* total_pairs"), ctypes.POINTER(ctypes.c_int))
    assert _compatible(_ctype_of("wealy_eval_plan** plan"), ctypes.POINTER(ctypes.c_void_p))
