"""-m gpu: masked reductions and distance_tensor_redux (CUDA kernel behind wealy_masked_reduce) vs
the reference's own outputs (tests/golden/masked.npz, redux.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _wt():
    from wealy_b200 import tensor_ops as wt
    return wt


def _close(a, b, tol=2e-6):
    a = a.detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    assert torch.equal(nan_a, nan_b)
    fin = ~nan_a
    assert torch.equal(torch.isinf(a), torch.isinf(b))
    fin &= ~torch.isinf(a)
    assert (a[fin] - b[fin]).abs().max().item() <= tol * max(1.0, b[fin].abs().max().item()) if fin.any() else True


def test_masked_reductions_against_reference(golden):
    G = golden("masked.npz")
    wt = _wt()
    x, mask = torch.from_numpy(G["x"]).cuda(), torch.from_numpy(G["mask"]).cuda()
    for fn in ("msum", "mmean", "mmin", "mmax"):
        f = getattr(wt, fn)
        _close(f(x, mask=mask), G[f"{fn}_all"])
        _close(f(x), G[f"{fn}_nomask"])
        _close(f(x, mask=mask, dim=2), G[f"{fn}_d2"])
        _close(f(x, mask=mask, dim=(1, 2), keepdim=True), G[f"{fn}_d12k"])
    _close(wt.mbest(x, 3, mask=mask, dim=-1), G["mbest_k3"])
    _close(wt.mworst(x, 3, mask=mask, dim=-1), G["mworst_k3"])
    q = wt.mmean(torch.tensor([1., 2., 3., 4.]).cuda(), mask=torch.tensor([True, False, False, False]).cuda())
    assert float(q) == 3.0                                 # mask True = excluded


@pytest.mark.parametrize("redux", ["min", "max", "mean", "minmean", "meanmin", "best", "best-3", "worst", "worst-2",
                                   "bestmin", "bestmin-2", "smin", "smeanmin", "sbest-4"])
def test_redux_against_reference(golden, redux):
    G = golden("redux.npz")
    wt = _wt()
    dist, mask = torch.from_numpy(G["dist"]).cuda(), torch.from_numpy(G["mask"]).cuda()
    _close(wt.distance_tensor_redux(dist, redux, mask=mask), G[f"m_{redux}"])
    _close(wt.distance_tensor_redux(dist, redux), G[f"n_{redux}"])


def test_redux_random_strategies_are_valid():
    # randmin / bpwr draw from the CUDA generator (the reference fixture used the CPU one), so check
    # the defining properties instead of the draws
    wt = _wt()
    g = torch.Generator().manual_seed(3)
    dist = (torch.rand(3, 4, 5, 6, generator=g) * 2).cuda()
    rowmin = dist.min(dim=-1)[0]
    r = wt.distance_tensor_redux(dist, "randmin")
    assert r.shape == (3, 4)
    assert bool(((r[..., None] - rowmin).abs().min(dim=-1)[0] < 1e-6).all())    # one of the per-chunk minima
    b = wt.distance_tensor_redux(dist, "bpwr")
    assert bool((b >= dist.flatten(2).min(dim=-1)[0] - 1e-6).all()) and bool((b <= dist.flatten(2).max(dim=-1)[0]).all())
    b1 = wt.distance_tensor_redux(dist, "bpwr-1")
    assert (b1 - dist.flatten(2).min(dim=-1)[0]).abs().max() < 1e-5            # one round = the global best pair
    with pytest.raises(NotImplementedError):
        wt.distance_tensor_redux(dist, "nope")


def test_long_rows_and_half_dtypes():
    wt = _wt()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 70000, generator=g).cuda()
    m = (torch.rand(3, 70000, generator=g) < 0.5).cuda()
    ref = torch.where(m, torch.zeros_like(x), x).double().sum(dim=1) / (~m).double().sum(dim=1)
    assert (wt.mmean(x, mask=m, dim=1).double() - ref).abs().max() < 1e-5
    assert float(wt.mmin(x, mask=m)) == float(torch.where(m, torch.full_like(x, float("inf")), x).min())
    xb = x[:, :1000].bfloat16()
    out = wt.mmax(xb, mask=m[:, :1000], dim=1)
    assert out.dtype == torch.bfloat16
    assert torch.equal(out, torch.where(m[:, :1000], torch.full_like(xb, -float("inf")), xb).max(dim=1)[0])


def test_float64_masked_reductions():
    """float64 inputs keep their dtype (double accumulators), like the reference's torch ops."""
    from wealy_b200 import tensor_ops as wt
    g = torch.Generator().manual_seed(31)
    x = torch.randn(37, 300, generator=g, dtype=torch.float64)
    m = torch.rand(37, 300, generator=g) < 0.3
    xc, mc = x.cuda(), m.cuda()
    inc = (~m).double()
    assert wt.msum(xc, mask=mc, dim=1).dtype == torch.float64
    assert (wt.msum(xc, mask=mc, dim=1).cpu() - (x * inc).sum(1)).abs().max() < 1e-12
    assert (wt.mmean(xc, mask=mc, dim=1).cpu() - (x * inc).sum(1) / inc.sum(1)).abs().max() < 1e-12
    assert torch.equal(wt.mmin(xc, mask=mc, dim=1).cpu(), torch.where(m, torch.full_like(x, float("inf")), x).min(1).values)
    assert torch.equal(wt.mmax(xc, mask=mc, dim=1).cpu(), torch.where(m, torch.full_like(x, float("-inf")), x).max(1).values)


@pytest.mark.parametrize("redux", ["min", "max", "mean", "minmean", "meanmin", "best", "best-3", "best-100", "worst-2", "bpwr",
                                   "bpwr-2", "smin", "smeanmin", "sminmean", "sbest-4", "sbpwr", "bestmin-2"])
@pytest.mark.parametrize("shape", [(3, 4, 5, 6), (2, 3, 7, 2), (5, 1, 1, 9), (2, 2, 16, 16), (1, 3, 32, 32)])
def test_fused_redux_equals_the_composition(redux, shape):
    """wealy_distance_redux (one launch, one warp per track pair) against the branch-for-branch composition of masked
    reductions (`fused=False`), with and without a mask that also empties whole rows, columns and blocks."""
    wt = _wt()
    g = torch.Generator().manual_seed(sum(shape) + len(redux))
    dist = (torch.rand(*shape, generator=g) * 2).cuda()
    mask = (torch.rand(*shape, generator=g) < 0.35).cuda()
    mask[0, 0, 0, :] = True                                    # a fully excluded row
    mask[-1, -1, :, -1] = True                                 # ... column
    if shape[0] > 1:
        mask[1, 0] = True                                      # ... block
    tol = 1e-5 if "bpwr" in redux else 2e-6                    # (the tie jitter of bpwr is drawn anew in each call)
    for m in (None, mask):
        a = wt.distance_tensor_redux(dist, redux, mask=m)
        b = wt.distance_tensor_redux(dist, redux, mask=m, fused=False)
        assert a.shape == b.shape == shape[:2]
        _close(a, b, tol=tol)
    a = wt.distance_tensor_redux(dist, redux, mask=mask, squeeze=False)
    assert a.shape == shape[:2] + (1, 1)
    _close(wt.distance_tensor_redux(dist.double(), redux, mask=mask), wt.distance_tensor_redux(dist, redux, mask=mask).double(), tol=1e-5)
    h = wt.distance_tensor_redux(dist.half(), redux, mask=mask)
    assert h.dtype == torch.float16
