"""-m gpu: MeanPool, the avg-pool collate and TripletLoss through the C ABI vs the reference's own outputs
(tests/golden/pooling_triplet.npz) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import losses as ol
from oracle import pooling as op

pytestmark = pytest.mark.gpu

T_CASES = ["t_a", "t_b", "t_single", "t_nopos"]
T_CFG = {"def": {}, "swap": {"swap": True, "margin": 0.5}, "p1sum": {"p": 1, "reduction": "sum"},
         "p3": {"p": 3, "margin": 1.0}}


def test_mean_pool_against_reference_outputs(golden):
    from wealy_b200.layers import MeanPool
    G = golden("pooling_triplet.npz")
    x, mask = torch.from_numpy(G["mp_x"]).cuda(), torch.from_numpy(G["mp_mask"]).cuda()
    for tag, m in (("masked", mask), ("plain", None)):
        xx = x.clone().requires_grad_(True)
        y = MeanPool()(xx, m)
        (y * torch.from_numpy(G[f"mp_{tag}_w"]).cuda()).sum().backward()
        assert np.allclose(y.detach().cpu().numpy(), G[f"mp_{tag}_y"], rtol=1e-5, atol=1e-6)
        assert np.allclose(xx.grad.cpu().numpy(), G[f"mp_{tag}_gx"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_mean_pool_large_and_half(dtype):
    from wealy_b200.layers import MeanPool
    g = torch.Generator().manual_seed(1)
    x = torch.randn(9, 130, 1501, generator=g)
    lens = torch.randint(1, 1501, (9,), generator=g)
    mask = torch.arange(1501)[None, :] < lens[:, None]
    want = op.mean_pool(x.to(dtype).float(), mask)
    got = MeanPool()(x.cuda().to(dtype), mask.cuda())
    assert got.dtype == dtype
    tol = 1e-5 if dtype == torch.float32 else (2e-3 if dtype == torch.float16 else 1e-2)
    assert torch.allclose(got.float().cpu(), want, rtol=tol, atol=tol)


def test_avg_pool_tracks_feeds_the_evaluator():
    from wealy_b200.layers import avg_pool_tracks
    g = torch.Generator().manual_seed(2)
    frames = [torch.randn(int(t), 96, generator=g).half() for t in torch.randint(1, 40, (57,), generator=g)]
    frames[5] = frames[5][:1]
    frames[9] = frames[9][:0]
    want, _ = op.avg_pool_tracks(frames, 96)
    got = avg_pool_tracks([f.cuda() for f in frames])
    assert got.dtype == torch.float32 and got.shape == (57, 96)
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6)
    # concatenated + offsets form
    lens = torch.tensor([0] + [f.shape[0] for f in frames])
    got2 = avg_pool_tracks(torch.cat(frames).cuda(), torch.cumsum(lens, 0))
    assert torch.equal(got, got2)


@pytest.mark.parametrize("name", T_CASES)
@pytest.mark.parametrize("tag", list(T_CFG))
def test_triplet_against_reference_outputs(golden, name, tag):
    from wealy_b200.losses import TripletLoss
    G = golden("pooling_triplet.npz")
    z = torch.from_numpy(G[f"{name}_z"]).cuda().requires_grad_(True)
    lab, idx = torch.from_numpy(G[f"{name}_label"]).cuda().clone(), torch.from_numpy(G[f"{name}_idx"]).cuda()
    mod = TripletLoss(**T_CFG[tag])
    loss, logd = mod(lab, idx, z)
    assert np.array_equal(lab.cpu().numpy(), G[f"{name}_{tag}_label_after"])
    a, p, n = mod._create_triplets(lab, idx)
    assert np.array_equal(a.cpu().numpy(), G[f"{name}_{tag}_anchors"])
    assert np.array_equal(p.cpu().numpy(), G[f"{name}_{tag}_pos"]) and np.array_equal(n.cpu().numpy(), G[f"{name}_{tag}_neg"])
    ref = float(G[f"{name}_{tag}_loss"])
    assert abs(float(loss) - ref) <= 1e-3 * abs(ref) + 1e-7          # north_star: loss within 1e-3 relative
    assert int("n_triplets" in logd) == int(G[f"{name}_{tag}_ntrip_key"])
    loss.backward()
    ref_g = torch.from_numpy(G[f"{name}_{tag}_grad"]).double()
    got_g = torch.zeros_like(ref_g) if z.grad is None else z.grad.cpu().double()
    assert float((got_g - ref_g).norm()) <= 1e-5 * float(ref_g.norm()) + 1e-9
    for k in ("v_zmax", "v_zmean", "v_zstd"):
        assert float(logd[k]) == pytest.approx(float(G[f"{name}_{tag}_log_{k}"]), rel=1e-5)


def test_triplet_none_reduction_and_big_batch():
    from wealy_b200.losses import TripletLoss
    g = torch.Generator().manual_seed(5)
    B, D = 4096, 512
    z = torch.randn(B, D, generator=g)
    lab = torch.randint(0, 900, (B,), generator=g)
    idx = torch.arange(B)
    zr = z.clone().requires_grad_(True)
    want, _ = ol.triplet(lab.clone(), idx, zr, reduction="none")
    zc = z.cuda().requires_grad_(True)
    got, _ = TripletLoss(reduction="none")(lab.cuda(), idx.cuda(), zc)
    assert got.shape == want.shape
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)
    w = torch.randn(want.shape, generator=g)
    (want * w).sum().backward()
    (got * w.cuda()).sum().backward()
    assert float((zc.grad.cpu() - zr.grad).norm()) <= 1e-5 * float(zr.grad.norm())
