"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/lib/tensor_ops.py, lib/losses.py) in the build container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small .npz
fixtures; tests compare the oracle (and, under -m gpu, the CUDA path) against them.
The ranking / AP / MR1 stage has no reference implementation: its
known-answer cases are hand-computed in tests/test_oracle_evaluator.py instead.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402

tops, rlosses = ref_import.load()

MODES = ("fro", "nfro", "euc", "neuc", "sqeuc", "nsqeuc", "cos", "cossim", "dot", "dotsim")


def np32(t):
    return t.detach().to(torch.float64).numpy() if t.dtype in (torch.float64,) else t.detach().float().numpy()


def gen_similarity():
    out = {}
    g = torch.Generator().manual_seed(1234)
    cases = {"a": (37, 53, 64), "b": (130, 257, 200), "c": (5, 3, 1024)}
    for name, (n, m, d) in cases.items():
        x = torch.randn(n, d, generator=g) * torch.rand(n, 1, generator=g).mul(3).add(0.1)
        y = torch.randn(m, d, generator=g) * 2.0
        if name == "a":
            x[3] = 0.0          # zero vector: cossim must be 0, not NaN (tensor_ops.py:169)
            y[5] = x[7]         # exact duplicate
        out[f"{name}_x"], out[f"{name}_y"] = x.numpy(), y.numpy()
        for mode in MODES:
            out[f"{name}_{mode}"] = tops.pairwise_distance_matrix(x, y, mode=mode).numpy()
    # 1-D inputs are treated as n x 1 column vectors (tensor_ops.py:154-156)
    v, w = torch.randn(6, generator=g), torch.randn(4, generator=g)
    out["vec_x"], out["vec_y"] = v.numpy(), w.numpy()
    for mode in MODES:
        out[f"vec_{mode}"] = tops.pairwise_distance_matrix(v, w, mode=mode).numpy()
    # output dtype follows the input dtype
    xb = torch.randn(16, 32, generator=g)
    out["dt_x"] = xb.numpy()
    for dt, tag in ((torch.bfloat16, "bf16"), (torch.float16, "f16"), (torch.float64, "f64")):
        r = tops.pairwise_distance_matrix(xb.to(dt), xb.to(dt), mode="cossim")
        assert r.dtype == dt
        out[f"dt_{tag}"] = r.double().numpy()
    np.savez_compressed(os.path.join(HERE, "sim_modes.npz"), **out)


def gen_masked():
    out = {}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 5, 6, generator=g)
    mask = torch.rand(4, 5, 6, generator=g) < 0.35
    mask[1, 2] = True  # a fully excluded row
    out["x"], out["mask"] = x.numpy(), mask.numpy()
    for fn in ("msum", "mmean", "mmin", "mmax"):
        f = getattr(tops, fn)
        out[f"{fn}_all"] = f(x, mask=mask).numpy()
        out[f"{fn}_nomask"] = f(x).numpy()
        out[f"{fn}_d2"] = f(x, mask=mask, dim=2).numpy()
        out[f"{fn}_d12k"] = f(x, mask=mask, dim=(1, 2), keepdim=True).numpy()
    out["mbest_k3"] = tops.mbest(x, 3, mask=mask, dim=-1).numpy()
    out["mworst_k3"] = tops.mworst(x, 3, mask=mask, dim=-1).numpy()   # quirk: always 0
    out["mmean_quirk"] = tops.mmean(torch.tensor([1., 2., 3., 4.]),
                                    mask=torch.tensor([True, False, False, False])).numpy()
    np.savez_compressed(os.path.join(HERE, "masked.npz"), **out)


def gen_redux():
    out = {}
    g = torch.Generator().manual_seed(11)
    dist = torch.rand(3, 4, 5, 6, generator=g) * 2
    mask = torch.rand(3, 4, 5, 6, generator=g) < 0.25
    out["dist"], out["mask"] = dist.numpy(), mask.numpy()
    for redux in ("min", "max", "mean", "minmean", "meanmin", "best", "best-3", "worst", "worst-2",
                  "bestmin", "bestmin-2", "smin", "smeanmin", "sbest-4"):
        out[f"m_{redux}"] = tops.distance_tensor_redux(dist, redux, mask=mask).numpy()
        out[f"n_{redux}"] = tops.distance_tensor_redux(dist, redux).numpy()
    for redux in ("randmin", "bpwr", "bpwr-2"):     # consume torch's global RNG
        torch.manual_seed(99)
        out[f"m_{redux}"] = tops.distance_tensor_redux(dist, redux, mask=mask).numpy()
        torch.manual_seed(99)
        out[f"n_{redux}"] = tops.distance_tensor_redux(dist, redux).numpy()
    np.savez_compressed(os.path.join(HERE, "redux.npz"), **out)


def gen_redux_ragged():
    """Ragged tracks: chunk-level cosine distances of [N, s, D] tracks whose last chunks are padding, reduced by the
    reference with the rectangular mask `query chunk invalid | candidate chunk invalid` -- the case the fused
    evaluation implements (wealy_eval_run_ragged)."""
    out = {}
    g = torch.Generator().manual_seed(23)
    n, m, s, d = 9, 11, 4, 16
    zq, zc = torch.randn(n, s, d, generator=g), torch.randn(m, s, d, generator=g)
    lq, lc = torch.randint(1, s + 1, (n,), generator=g), torch.randint(1, s + 1, (m,), generator=g)
    lq[0], lc[0] = 1, s
    dist = tops.pairwise_distance_matrix(zq.reshape(n * s, d), zc.reshape(m * s, d), mode="cos")
    dist = dist.reshape(n, s, m, s).permute(0, 2, 1, 3).contiguous()
    mask = (torch.arange(s)[None, :] >= lq[:, None])[:, None, :, None] | (torch.arange(s)[None, :] >= lc[:, None])[None, :, None, :]
    out["zq"], out["zc"], out["lq"], out["lc"] = zq.numpy(), zc.numpy(), lq.numpy(), lc.numpy()
    for redux in ("min", "max", "mean", "minmean", "meanmin"):
        out[f"r_{redux}"] = tops.distance_tensor_redux(dist, redux, mask=mask).numpy()
    np.savez_compressed(os.path.join(HERE, "redux_ragged.npz"), **out)


def _loss_case(B, D, dtype, seed, per_clique=4, dup_idx=True, single_label=False):
    g = torch.Generator().manual_seed(seed)
    z = (torch.randn(B, D, generator=g) * 1.5 + 0.1).to(dtype)
    lab = torch.arange(B) // per_clique
    if single_label:
        lab = torch.zeros(B, dtype=torch.long)
    idx = torch.arange(B)
    if dup_idx:
        idx[3] = idx[2]
        idx[B - 1] = idx[B - 2]
    perm = torch.randperm(B, generator=g)
    return z, lab[perm].clone(), idx[perm].clone()


def gen_losses():
    out = {}
    cases = {
        "f64": dict(B=96, D=64, dtype=torch.float64, seed=3),
        "f32": dict(B=256, D=128, dtype=torch.float32, seed=4),
        "f32_big": dict(B=256, D=1024, dtype=torch.float32, seed=5, per_clique=8),
        "nopos": dict(B=64, D=32, dtype=torch.float32, seed=6, per_clique=1),       # no positives at all
        "single": dict(B=200, D=32, dtype=torch.float32, seed=8, single_label=True),  # label-noise path
    }
    for name, kw in cases.items():
        z, lab, idx = _loss_case(**kw)
        keep = (lambda t: t.double().numpy()) if z.dtype == torch.float64 else (lambda t: t.float().numpy())
        sub = 4 if name == "f32_big" else 1          # big case: keep every 4th gradient row only
        out[f"{name}_z"] = keep(z)
        out[f"{name}_gradrows"] = np.arange(0, z.shape[0], sub)
        out[f"{name}_label"], out[f"{name}_idx"] = lab.numpy(), idx.numpy()
        for tag, mod, extra in (
            ("ntx", rlosses.NTXentLoss(temperature=0.1), None),
            ("ntx05", rlosses.NTXentLoss(temperature=0.5), None),
            ("clews", rlosses.CLEWSLoss(), None),
            ("clews_step", rlosses.CLEWSLoss(gamma=6.0, b=0.5, uniformity_weight=0.8, warmup_steps=100),
             {"global_step": 9}),
        ):
            zz = z.clone().requires_grad_(True)
            lab_in = lab.clone()
            loss, logd = mod(lab_in, idx.clone(), zz, extra=extra)
            (grad,) = torch.autograd.grad(loss, zz, allow_unused=True)
            out[f"{name}_{tag}_loss"] = loss.detach().double().numpy()
            out[f"{name}_{tag}_grad"] = keep((torch.zeros_like(zz) if grad is None else grad).detach()[::sub])
            out[f"{name}_{tag}_label_after"] = lab_in.numpy()
            for k, v in logd.items():
                out[f"{name}_{tag}_log_{k}"] = torch.as_tensor(v).detach().double().numpy()
        # the non-"numerically friendly" CLEWS branch (losses.py:245)
        zz = z.clone().requires_grad_(True)
        loss, _ = rlosses.CLEWSLoss()(lab.clone(), idx.clone(), zz, numerically_friendly=False)
        (grad,) = torch.autograd.grad(loss, zz)
        out[f"{name}_clews_nf_loss"] = loss.detach().double().numpy()
        out[f"{name}_clews_nf_grad"] = keep(grad.detach()[::sub])
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)


def gen_pooling_triplet():
    """f3 / f4: the reference's MeanPool (lib/layers.py:6-30) and TripletLoss (lib/losses.py:76-171), unmodified."""
    import importlib
    layers = importlib.import_module("lib.layers")
    out = {}
    g = torch.Generator().manual_seed(21)
    x = torch.randn(5, 12, 37, generator=g)
    mask = torch.rand(5, 37, generator=g) < 0.7
    mask[2] = False                                   # a fully padded item: 0 / 1e-8
    mask[3] = True
    out["mp_x"], out["mp_mask"] = x.numpy(), mask.numpy()
    mp = layers.MeanPool()
    for tag, m in (("masked", mask), ("plain", None)):
        xx = x.clone().requires_grad_(True)
        y = mp(xx, m)
        w = torch.randn(y.shape, generator=torch.Generator().manual_seed(3))
        (gx,) = torch.autograd.grad((y * w).sum(), xx)
        out[f"mp_{tag}_y"], out[f"mp_{tag}_w"], out[f"mp_{tag}_gx"] = y.detach().numpy(), w.numpy(), gx.numpy()
    # triplet loss: clique batches incl. duplicates of idx, a clique of one (no positive) and a single-label batch
    for name, (B, D, per, single) in {"t_a": (48, 64, 4, False), "t_b": (33, 40, 3, False), "t_single": (24, 16, 24, True),
                                      "t_nopos": (12, 8, 1, False)}.items():
        gg = torch.Generator().manual_seed(len(name) * 7 + B)
        z = torch.randn(B, D, generator=gg) * 0.7
        lab = (torch.arange(B) // per).long()
        lab = lab[torch.randperm(B, generator=gg)]
        if single:
            lab[:] = 5
        idx = torch.arange(B).long()
        if B > 20:
            idx[7] = idx[3]                            # duplicate version id
        out[f"{name}_z"], out[f"{name}_label"], out[f"{name}_idx"] = z.numpy(), lab.numpy(), idx.numpy()
        for tag, kw in (("def", {}), ("swap", {"swap": True, "margin": 0.5}), ("p1sum", {"p": 1, "reduction": "sum"}),
                        ("p3", {"p": 3, "margin": 1.0})):
            mod = rlosses.TripletLoss(**kw)
            zz = z.clone().requires_grad_(True)
            lab_in = lab.clone()
            loss, logd = mod(lab_in, idx.clone(), zz)
            (grad,) = torch.autograd.grad(loss, zz, allow_unused=True)
            out[f"{name}_{tag}_loss"] = loss.detach().double().numpy()
            out[f"{name}_{tag}_grad"] = (torch.zeros_like(zz) if grad is None else grad).numpy()
            out[f"{name}_{tag}_label_after"] = lab_in.numpy()
            a, pp, nn_ = mod._create_triplets(lab_in, idx)
            out[f"{name}_{tag}_anchors"], out[f"{name}_{tag}_pos"], out[f"{name}_{tag}_neg"] = (
                a.long().numpy(), pp.long().numpy(), nn_.long().numpy())
            out[f"{name}_{tag}_ntrip_key"] = np.array(int("n_triplets" in logd))
            for k in ("v_zmax", "v_zmean", "v_zstd"):
                out[f"{name}_{tag}_log_{k}"] = logd[k].detach().double().numpy()
    np.savez_compressed(os.path.join(HERE, "pooling_triplet.npz"), **out)


if __name__ == "__main__":
    gen_similarity()
    gen_masked()
    gen_redux()
    gen_redux_ragged()
    gen_losses()
    gen_pooling_triplet()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
