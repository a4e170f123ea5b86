"""-m gpu: fused NT-Xent / CLEWS forward + backward (through the C ABI) vs the reference's own
outputs (tests/golden/losses.npz) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import losses as ol

pytestmark = pytest.mark.gpu

CASES = ["f64", "f32", "f32_big", "nopos", "single"]
LOSS_RTOL = 1e-3          # north_star: loss within 1e-3 relative
GRAD_RTOL_FP32 = 1e-5     # rel-L2 <= 1e-5 (fp32-grade mode)
GRAD_RTOL_FAST = 1e-3     # single-pass fp16 mode


def _wl():
    from wealy_b200 import losses as wl
    return wl


def _mods(precision):
    wl = _wl()
    return (
        ("ntx", wl.NTXentLoss(0.1, precision=precision), None, {}),
        ("ntx05", wl.NTXentLoss(0.5, precision=precision), None, {}),
        ("clews", wl.CLEWSLoss(precision=precision), None, {}),
        ("clews_step", wl.CLEWSLoss(gamma=6.0, b=0.5, uniformity_weight=0.8, warmup_steps=100, precision=precision),
         {"global_step": 9}, {}),
        ("clews_nf", wl.CLEWSLoss(precision=precision), None, {"numerically_friendly": False}),
    )


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
def test_against_reference_outputs(golden, name, precision):
    G = golden("losses.npz")
    z = torch.from_numpy(G[f"{name}_z"]).float()
    lab, idx = torch.from_numpy(G[f"{name}_label"]), torch.from_numpy(G[f"{name}_idx"])
    rows = torch.from_numpy(G[f"{name}_gradrows"])
    gtol = GRAD_RTOL_FP32 if precision == "fp16x3" else GRAD_RTOL_FAST
    for tag, mod, extra, kw in _mods(precision):
        zz = z.cuda().requires_grad_(True)
        lab_c = lab.cuda().clone()
        loss, logd = mod(lab_c, idx.cuda(), zz, extra=extra, **kw)
        loss.backward()
        torch.cuda.synchronize()
        ref_l = float(G[f"{name}_{tag}_loss"])
        assert abs(float(loss.detach()) - ref_l) <= LOSS_RTOL * abs(ref_l) + 1e-7, (tag, float(loss), ref_l)
        ref_g = torch.from_numpy(G[f"{name}_{tag}_grad"]).double()
        g = zz.grad.cpu().double()[rows]
        # gradients: relative L2 against the reference's autograd; rows whose reference gradient is
        # (numerically) nothing are held to an absolute bar instead
        denom = float(ref_g.norm())
        if denom > 1e-4:
            # the f64 fixture was produced in double; our arithmetic is fp32-grade
            bar = gtol * (20 if name in ("nopos",) else 1) * (2 if name == "f64" else 1)
            assert float((g - ref_g).norm()) <= bar * denom, (tag, float((g - ref_g).norm()) / denom)
        else:
            assert float((g - ref_g).abs().max()) <= 1e-6
        if tag in ("ntx", "clews", "clews_step"):
            assert np.array_equal(lab_c.cpu().numpy(), G[f"{name}_{tag}_label_after"])   # in-place label noise
            ref_keys = sorted(k.split("_log_")[1] for k in G.files if k.startswith(f"{name}_{tag}_log_"))
            assert sorted(logd.keys()) == ref_keys
            for k, v in logd.items():
                ref_v = float(G[f"{name}_{tag}_log_{k}"])
                assert abs(float(v) - ref_v) <= 1e-3 * max(1.0, abs(ref_v)), (tag, k, float(v), ref_v)


def test_c4_shape_bf16_batch():
    """BASELINE configs[3]: batch 4096 x 1024 bf16, 4 items per clique."""
    from wealy_b200.data import synth
    wl = _wl()
    s = synth.make_loss_batch(4096, 1024, seed=0, dtype=torch.bfloat16)
    for make, oracle in ((lambda **kw: wl.NTXentLoss(0.1, **kw), lambda l, i, z: ol.ntxent(l, i, z, 0.1)),
                         (lambda **kw: wl.CLEWSLoss(**kw), lambda l, i, z: ol.clews(l, i, z))):
        # default: the loss carries z's dtype, like the reference (lib/losses.py:65-66 builds it from z's own dtype)
        l16, logd16 = make()(s["label"].cuda(), s["idx"].cuda(), s["z"].cuda())
        assert l16.dtype == torch.bfloat16 and all(v.dtype == torch.bfloat16 for k, v in logd16.items() if k != "uniformity_weight")
        z = s["z"].cuda().requires_grad_(True)
        loss, logd = make(loss_dtype=torch.float32)(s["label"].cuda(), s["idx"].cuda(), z)     # unrounded value
        loss.backward()
        torch.cuda.synchronize()
        zr = s["z"].float().requires_grad_(True)          # the same bf16 values, reference arithmetic in fp32
        loss_o, _ = oracle(s["label"].clone(), s["idx"], zr)
        loss_o.backward()
        assert loss.dtype == torch.float32 and z.grad.dtype == torch.bfloat16
        assert abs(float(l16) - float(loss_o)) <= 2 ** -8 * abs(float(loss_o))      # one bf16 rounding of the fp32 result
        assert abs(float(loss.detach()) - float(loss_o)) <= LOSS_RTOL * abs(float(loss_o))
        rel = float((z.grad.cpu().float() - zr.grad).norm() / zr.grad.norm())
        assert rel <= 4e-3, rel                           # bf16 output rounding: 2^-9 per element


def test_upstream_gradient_and_3d_input():
    from wealy_b200.data import synth
    wl = _wl()
    s = synth.make_loss_batch(192, 96, seed=1)
    z = s["z"].cuda().requires_grad_(True)
    loss, _ = wl.CLEWSLoss()(s["label"].cuda(), s["idx"].cuda(), z.unsqueeze(1))      # (B, 1, C) accepted
    (3.0 * loss).backward()
    g3 = z.grad.clone()
    z.grad = None
    loss2, _ = wl.CLEWSLoss()(s["label"].cuda(), s["idx"].cuda(), z)
    loss2.backward()
    assert torch.allclose(g3, 3.0 * z.grad, rtol=1e-5, atol=1e-9)
    assert abs(float(loss.detach()) - float(loss2.detach())) < 1e-7


def test_global_step_attribute_and_warmup():
    from wealy_b200.data import synth
    wl = _wl()
    s = synth.make_loss_batch(128, 64, seed=2)
    lab, idx, z = s["label"].cuda(), s["idx"].cuda(), s["z"].cuda()
    m = wl.CLEWSLoss(uniformity_weight=0.5, warmup_steps=1000)
    _, d0 = m(lab, idx, z)
    assert abs(float(d0["uniformity_weight"]) - 0.5) < 1e-9               # no step supplied -> full weight
    m.global_step = 99
    _, d1 = m(lab, idx, z)
    assert abs(float(d1["uniformity_weight"]) - 0.05) < 1e-9              # lib/losses.py:255-258
    _, d2 = m(lab, idx, z, extra={"global_step": 499})
    assert abs(float(d2["uniformity_weight"]) - 0.25) < 1e-9
    lo, _ = ol.clews(s["label"].clone(), s["idx"], s["z"], step=499)
    assert abs(float(d2["l_main"]) - float(lo)) <= 1e-5 * abs(float(lo))


def test_ragged_batch_sizes():
    wl = _wl()
    for b, d in ((4, 3), (33, 17), (130, 70), (257, 129)):
        g = torch.Generator().manual_seed(b)
        z = torch.randn(b, d, generator=g)
        lab = torch.arange(b) // 2
        idx = torch.arange(b)
        for mod, fn in ((wl.NTXentLoss(0.2), lambda l, i, zz: ol.ntxent(l, i, zz, 0.2)),
                        (wl.CLEWSLoss(), lambda l, i, zz: ol.clews(l, i, zz))):
            zc = z.cuda().requires_grad_(True)
            loss, _ = mod(lab.cuda(), idx.cuda(), zc)
            loss.backward()
            zr = z.clone().requires_grad_(True)
            lo, _ = fn(lab.clone(), idx, zr)
            lo.backward()
            assert abs(float(loss.detach()) - float(lo)) <= 1e-5 * max(1.0, abs(float(lo)))
            assert float((zc.grad.cpu() - zr.grad).norm()) <= 2e-5 * float(zr.grad.norm()) + 1e-9


@pytest.mark.parametrize("kind", ["ntxent", "clews"])
def test_forward_backward_is_cuda_graph_capturable(kind):
    """The loss step is launch bound (about ten short kernels): it must be capturable in a CUDA graph -- no host
    synchronisation, no pageable host copies anywhere between the call and z.grad."""
    from wealy_b200.data import synth
    wl = _wl()
    s = synth.make_loss_batch(512, 256, seed=2, dtype=torch.bfloat16, device="cuda")
    mod = wl.NTXentLoss(0.1) if kind == "ntxent" else wl.CLEWSLoss()
    z = s["z"].clone().requires_grad_(True)
    lab, idx = s["label"].clone(), s["idx"].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            z.grad = None
            loss, _ = mod(lab, idx, z)
            loss.backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref_loss, ref_grad = float(loss), z.grad.clone()
    g = torch.cuda.CUDAGraph()
    z.grad = None
    with torch.cuda.graph(g):
        loss_g, _ = mod(lab, idx, z)
        loss_g.backward()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert float(loss_g) == ref_loss and torch.equal(z.grad, ref_grad)


def test_float64_batch_and_id_overflow():
    """float64 batches: rounded to float32 on entry, loss / gradient come back as float64 (the reference returns the
    input dtype).  Labels / ids beyond 32 bits would be truncated by the packed id records: the loss is NaN, loudly."""
    from wealy_b200.data import synth
    wl = _wl()
    s = synth.make_loss_batch(256, 64, seed=4)
    for mod, oracle in ((wl.NTXentLoss(0.1), lambda l, i, z: ol.ntxent(l, i, z, 0.1)), (wl.CLEWSLoss(), lambda l, i, z: ol.clews(l, i, z))):
        z = s["z"].double().cuda().requires_grad_(True)
        loss, _ = mod(s["label"].cuda(), s["idx"].cuda(), z)
        loss.backward()
        zr = s["z"].double().requires_grad_(True)
        lo, _ = oracle(s["label"].clone(), s["idx"], zr)
        lo.backward()
        assert loss.dtype == torch.float64 and z.grad.dtype == torch.float64
        assert abs(float(loss) - float(lo)) <= 1e-6 * abs(float(lo))
        assert float((z.grad.cpu() - zr.grad).norm()) <= 1e-5 * float(zr.grad.norm())
        big = s["label"].cuda() + (1 << 32)                       # equal low 32 bits would alias labels 2^32 apart
        bad, _ = mod(big, s["idx"].cuda(), s["z"].cuda())
        assert torch.isnan(bad)
        bad, _ = mod(s["label"].cuda(), s["idx"].cuda() - (1 << 33), s["z"].cuda())
        assert torch.isnan(bad)
