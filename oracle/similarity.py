"""Oracle restatement of the pairwise similarity / distance matrix (TEST INFRASTRUCTURE).

Follows /root/reference/lib/tensor_ops.py:
  * pairwise_distance_matrix            lines 152-176
  * pairwise_euclidean_distance_matrix  lines 131-149

Written against torch CPU tensors because the reference's arithmetic *is* torch's
(ATen matmul / cdist); the installed torch CPU build is the oracle arithmetic
(SURVEY.md section 8(c)).  Pinned by tests/golden/sim_*.npz.
"""
import torch

_SIM_MODES = ("cos", "cossim", "dot", "dotsim")
_CDIST_MODES = ("fro", "nfro", "euc", "neuc")
_SQ_MODES = ("sqeuc", "nsqeuc")
ALL_MODES = _CDIST_MODES + _SQ_MODES + _SIM_MODES


def squared_euclidean(x, y):
    """|x|^2 - 2 x.y + |y|^2 with non-positive entries clamped to zero.
    Reference: lib/tensor_ops.py:132-137."""
    xx = (x * x).sum(dim=1)[:, None]
    yy = (y * y).sum(dim=1)[None, :]
    out = xx - 2.0 * (x @ y.t()) + yy
    return torch.where(out <= 0, torch.zeros_like(out), out)


def euclidean(x, y, squared=False, eps=1e-6):
    """Reference: lib/tensor_ops.py:131-149 (safe sqrt: zeros stay exactly zero)."""
    d2 = squared_euclidean(x, y)
    if squared:
        return d2
    zero = d2 == 0
    root = torch.sqrt(torch.where(zero, torch.full_like(d2, eps), d2))
    return torch.where(zero, torch.zeros_like(root), root)


def l2_scale(x, eps=1e-6):
    """x / (|x|_2 + eps) row-wise.  Reference: lib/tensor_ops.py:169-170.
    NOTE: eps is *added to the norm* (not a clamp as in F.normalize)."""
    nrm = torch.linalg.vector_norm(x, ord=2, dim=-1, keepdim=True)
    return x / (nrm + eps)


def distance_matrix(x, y, mode="fro", p=2, eps=1e-6):
    """Reference: lib/tensor_ops.py:152-176.  Returns an (n, m) tensor in x's dtype."""
    if not (x.ndim == y.ndim and x.ndim <= 2):
        raise AssertionError("x and y must have the same rank <= 2")  # :153
    if x.ndim == 1:  # :154-156 -> column vectors (outer product)
        x, y = x[:, None], y[:, None]
    if mode in _CDIST_MODES:  # :157-162
        if mode in ("euc", "neuc"):
            p = 2
        out = torch.cdist(x[None], y[None], p=p)[0]
        if mode in ("nfro", "neuc"):
            out = out / (x.shape[-1] ** (1.0 / p))
        return out
    if mode in _SQ_MODES:  # :163-166
        out = squared_euclidean(x, y)
        return out / x.shape[-1] if mode == "nsqeuc" else out
    if mode in _SIM_MODES:  # :167-173
        if mode in ("cos", "cossim"):
            x, y = l2_scale(x, eps), l2_scale(y, eps)
        out = x @ y.t()
        return 1 - out if mode in ("cos", "dot") else out
    raise NotImplementedError(mode)  # :175
