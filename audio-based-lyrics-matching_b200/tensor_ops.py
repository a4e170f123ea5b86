"""Drop-in for the hot functions of /root/reference/lib/tensor_ops.py, computed on sm_100a.

  pairwise_distance_matrix(x, y, mode="fro", p=2, eps=1e-6)        lib/tensor_ops.py:152-176
  pairwise_euclidean_distance_matrix(x, y, squared=False, eps=1e-6) lib/tensor_ops.py:131-149

Same names, argument meaning, return shape / dtype and error behaviour as the reference
(AssertionError for rank violations, NotImplementedError for unknown modes).  The contraction runs
in the tcgen05 kernel behind wealy_sim_matrix (include/wealy_b200.h); `precision` selects the
tensor-core mode ("fp16x3": fp32-grade hi/lo split, the default; "fp16": one pass, ~1e-4 abs error).
Inputs must be CUDA tensors -- there is no CPU fallback.
"""
import ctypes
import os

import torch

from . import _native as N

_PRECISIONS = {"fp16x3": 3, "fp16": 1}
_default_precision = os.environ.get("WEALY_PRECISION", "fp16x3")


def set_default_precision(name):
    global _default_precision
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    _default_precision = name


def passes_of(precision=None):
    name = _default_precision if precision is None else precision
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    return _PRECISIONS[name]


def _rows(t):
    """2-D view with unit inner stride (the C ABI takes a row stride, not arbitrary strides)."""
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        t = t.contiguous()
    if t.shape[-1] == 1 and t.stride(0) < 1:
        t = t.contiguous()
    return t


def _sim_matrix(x, y, mode, eps, post, precision):
    N.require_cuda(x, y)
    if x.dtype != y.dtype:
        raise RuntimeError(f"expected x and y to have the same dtype, got {x.dtype} and {y.dtype}")
    if x.shape[1] != y.shape[1]:
        raise RuntimeError(f"size mismatch: x is {tuple(x.shape)}, y is {tuple(y.shape)}")
    code = N.dtype_code(x.dtype)
    n, d = x.shape
    m = y.shape[0]
    out = torch.empty((n, m), dtype=x.dtype, device=x.device)
    if n == 0 or m == 0:
        return out
    if d == 0:
        raise NotImplementedError("wealy_b200: zero-width embeddings")
    same = x is y
    x = _rows(x)
    y = x if same else _rows(y)
    passes = passes_of(precision)
    with torch.cuda.device(x.device):
        ws_bytes = N.lib.wealy_sim_matrix_workspace_bytes(n, m, d, passes)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        N.check(N.lib.wealy_sim_matrix(
            x.data_ptr(), n, x.stride(0), y.data_ptr(), m, y.stride(0), d, code, mode, float(eps), float(post),
            passes, out.data_ptr(), out.stride(0), code, ws.data_ptr(), ws_bytes, N.stream_ptr(x.device)))
    return out


def pairwise_euclidean_distance_matrix(x, y, squared=False, eps=1e-6, precision=None):
    """lib/tensor_ops.py:131-149: |x|^2 - 2 x.y + |y|^2, clamped at 0, optional sqrt (zeros stay 0).
    `eps` only guards the reference's autograd through sqrt(0); the forward value does not depend on it."""
    if x.requires_grad or y.requires_grad:
        raise NotImplementedError("wealy_b200.pairwise_euclidean_distance_matrix: autograd is not provided; "
                                  "use wealy_b200.losses for the fused differentiable losses")
    return _sim_matrix(x, y, N.MODE_SQEUC if squared else N.MODE_EUC, 0.0, 1.0, precision)


def pairwise_distance_matrix(x, y, mode="fro", p=2, eps=1e-6, precision=None):
    """lib/tensor_ops.py:152-176.  Returns an (n, m) tensor with x's dtype."""
    assert x.ndim == y.ndim and x.ndim <= 2
    if x.ndim == 1:  # :154-156 -- 1-D inputs are n x 1 column vectors
        x = x.unsqueeze(-1)
        y = y.unsqueeze(-1)
    if x.ndim == 0:
        raise NotImplementedError("wealy_b200: 0-d inputs")
    if x.requires_grad or y.requires_grad:
        raise NotImplementedError("wealy_b200.pairwise_distance_matrix: autograd is not provided; "
                                  "use wealy_b200.losses for the fused differentiable losses")
    if mode == "euc" or mode == "neuc":
        p = 2
    d = x.size(-1)
    if mode in ("fro", "nfro", "euc", "neuc"):
        if p != 2:
            raise NotImplementedError("wealy_b200: cdist modes are built for p=2 only (p-norms with p != 2 are "
                                      "not a contraction; out of the hot path)")
        post = 1.0 if mode in ("fro", "euc") else 1.0 / (d ** (1 / p))
        return _sim_matrix(x, y, N.MODE_EUC, 0.0, post, precision)
    if mode in ("sqeuc", "nsqeuc"):
        return _sim_matrix(x, y, N.MODE_SQEUC, 0.0, 1.0 if mode == "sqeuc" else 1.0 / d, precision)
    if mode in ("cos", "cossim", "dot", "dotsim"):
        code = {"cossim": N.MODE_COSSIM, "cos": N.MODE_COS, "dotsim": N.MODE_DOTSIM, "dot": N.MODE_DOT}[mode]
        return _sim_matrix(x, y, code, eps, 1.0, precision)
    raise NotImplementedError
