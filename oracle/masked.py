"""Oracle restatement of the masked reductions and the multi-chunk distance reduction
(TEST INFRASTRUCTURE).

Follows /root/reference/lib/tensor_ops.py:
  * msum 182-194, mmean 197-212, mmin 215-235, mmax 238-258, mrand 261-266,
    mbest 269-274, mworst 277-282
  * distance_tensor_redux 288-373

Mask polarity: ``mask == True`` means EXCLUDED (lines 186, 201, 219).
Reference quirks that are reproduced on purpose:
  * mworst builds its validity mask as ``x >= -inf`` which excludes everything, so it
    always returns 0 (line 282);
  * ``redux="bestmin..."`` is unreachable because ``startswith("best")`` (line 336) wins.
Pinned by tests/golden/masked_*.npz and redux_*.npz.
"""
import torch

INF = float("inf")


def _keep(x, mask):
    return torch.ones_like(x) if mask is None else (~mask).to(x.dtype)


def _pad_dims(t, ndim):
    while t.ndim < ndim:
        t = t[None]
    return t


def msum(x, mask=None, dim=None, keepdim=False):
    w = _keep(x, mask) * x
    if dim is None:
        s = w.sum()
        return _pad_dims(s, x.ndim) if keepdim else s
    return w.sum(dim=dim, keepdim=keepdim)


def mmean(x, mask=None, dim=None, keepdim=False, eps=1e-7):
    k = _keep(x, mask)
    if dim is None:
        num, den = (k * x).sum(), k.sum()
        if keepdim:
            num, den = _pad_dims(num, x.ndim), _pad_dims(den, x.ndim)
    else:
        num = (k * x).sum(dim=dim, keepdim=keepdim)
        den = k.sum(dim=dim, keepdim=keepdim)
    return num / den.clamp(min=eps)


def _extreme(x, mask, dim, keepdim, fill, take_min):
    t = x if mask is None else torch.where(mask, fill, x)
    if dim is None:
        t = t.min() if take_min else t.max()
        return _pad_dims(t, x.ndim) if keepdim else t
    dims = [dim] if isinstance(dim, int) else list(dim)
    for d in dims:  # one dim at a time, keepdim, as lines 230-231
        t = (t.min(d, keepdim=True) if take_min else t.max(d, keepdim=True))[0]
    if not keepdim:
        for d in dims:
            t = t.squeeze(d)
    return t


def mmin(x, mask=None, dim=None, keepdim=False, ctt=INF):
    return _extreme(x, mask, dim, keepdim, ctt, True)


def mmax(x, mask=None, dim=None, keepdim=False, ctt=-INF):
    return _extreme(x, mask, dim, keepdim, ctt, False)


def mrand(x, mask=None, dim=None, keepdim=False, ctt=INF, eps=1e-7):
    r = torch.rand_like(x)
    if mask is not None:
        r = torch.where(mask, ctt, r)
    not_chosen = r > mmin(r, mask=mask, dim=dim, keepdim=True, ctt=ctt)
    return mmean(x, mask=not_chosen, dim=dim, keepdim=keepdim, eps=eps)


def mbest(x, k, mask=None, dim=None, keepdim=False, ctt=INF, eps=1e-7):
    assert type(dim) == int
    if mask is not None:
        x = torch.where(mask, ctt, x)
    small = x.topk(k, dim=dim, largest=False)[0]
    return mmean(small, mask=small >= ctt, dim=dim, keepdim=keepdim, eps=eps)


def mworst(x, k, mask=None, dim=None, keepdim=False, ctt=-INF, eps=1e-7):
    assert type(dim) == int
    if mask is not None:
        x = torch.where(mask, ctt, x)
    big = x.topk(k, dim=dim, largest=True)[0]
    # quirk (line 282): with ctt = -inf the mask ``big >= ctt`` is all-True => result 0
    return mmean(big, mask=big >= ctt, dim=dim, keepdim=keepdim, eps=eps)


def _k_of(redux, limit):
    return 1 if "-" not in redux else max(1, min(int(redux.rsplit("-", 1)[-1]), limit))


def distance_tensor_redux(dist, redux, mask=None, squeeze=True, eps=1e-7, inf=1e12):
    """(b1, b2, s1, s2) chunk-level distances -> (b1, b2).  Reference: lines 288-373."""
    last2 = (-1, -2)
    if redux == "min":
        out = mmin(dist, mask, last2, True, inf)
    elif redux == "max":
        out = mmax(dist, mask, last2, True, -inf)
    elif redux == "mean":
        out = mmean(dist, mask, last2, True, eps)
    elif redux == "minmean":
        out = mmean(dist, mask, -1, True, eps)
        out = mmin(out, mask, last2, True, inf)
    elif redux == "meanmin":
        out = mmin(dist, mask, -1, True, inf)
        out = mmean(out, mask, last2, True, eps)
    elif redux == "randmin":
        out = mmin(dist, mask, -1, True, inf)
        out = mrand(out, mask, last2, True, inf, eps)
    elif redux.startswith("bpwr"):
        if dist.size(3) < dist.size(2):
            dist = dist.transpose(2, 3)
            mask = None if mask is None else mask.transpose(2, 3)
        n = dist.size(2) if "-" not in redux else _k_of(redux, dist.size(2))
        dist = dist + eps * torch.rand_like(dist)
        if mask is None:
            mask = dist > inf
        chosen = dist > inf
        for it in range(n):
            cur = mmin(dist, mask, last2, True, inf)
            chosen = chosen | ((dist <= cur) & ~mask)
            if it < n - 1:
                mask = (mask
                        | (mmin(dist, mask, -1, True, inf) <= cur)
                        | (mmin(dist, mask, -2, True, inf) <= cur))
        out = mmean(dist, ~chosen, last2, True, eps)
    elif redux.startswith("best"):  # also swallows "bestmin..." (quirk)
        k = _k_of(redux, dist.size(2) * dist.size(3))
        flat = dist.reshape(dist.size(0), dist.size(1), 1, -1)
        fmask = None if mask is None else mask.reshape(mask.size(0), mask.size(1), 1, -1)
        out = mbest(flat, k, fmask, -1, True, inf, eps)
    elif redux.startswith("worst"):
        k = _k_of(redux, dist.size(2) * dist.size(3))
        flat = dist.reshape(dist.size(0), dist.size(1), 1, -1)
        fmask = None if mask is None else mask.reshape(mask.size(0), mask.size(1), 1, -1)
        out = mworst(flat, k, fmask, -1, True, -inf, eps)
    elif redux[0] == "s":
        a = distance_tensor_redux(dist, redux[1:], mask, squeeze=False)
        b = distance_tensor_redux(dist.transpose(2, 3), redux[1:],
                                  None if mask is None else mask.transpose(2, 3), squeeze=False)
        out = 0.5 * (a + b.transpose(2, 3))
    else:
        raise NotImplementedError(redux)
    return out.squeeze((-1, -2)) if squeeze else out
