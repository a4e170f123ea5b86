"""GPU bring-up diagnostics (developer tool, run under gpurun): each stage in its own process so a
trap in one kernel does not poison the rest.  Usage: python tools/gpu_diag.py <stage> [args]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402
from wealy_b200 import tensor_ops as wt, evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402
from oracle import similarity as osim, evaluator as oev  # noqa: E402


def stage_sim(precision, n, m, d, mode="cossim"):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, d, generator=g) * 3
    y = torch.randn(m, d, generator=g)
    ref = osim.distance_matrix(x.double(), y.double(), mode=mode)
    out = wt.pairwise_distance_matrix(x.cuda(), y.cuda(), mode=mode, precision=precision)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs()
    print(f"sim {mode} {precision} n={n} m={m} d={d}: max_err={err.max().item():.3e} mean_err={err.mean().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3f}", flush=True)
    if err.max().item() > 1e-2:
        bad = (err > 1e-2).nonzero()
        print("  first bad entries:", bad[:8].tolist(), "count", bad.shape[0], flush=True)
        print("  out[0,:8]", out[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist(), flush=True)


def stage_eval(precision, n, d, topk=0):
    s = synth.make_eval_set(n, d, seed=0)
    t = time.time()
    aps_o, r1_o = oev.evaluate_rankcount(s["c"], s["i"], s["z"], s["c"], s["i"], s["z"]) if n <= 4000 else \
        oev.evaluate_argsort(s["c"], s["i"], s["z"], s["c"], s["i"], s["z"])
    t_or = time.time() - t
    zc, cc, ic = s["z"].cuda(), s["c"].cuda(), s["i"].cuda()
    res = we.evaluate(cc, ic, zc, cc, ic, zc, precision=precision, topk=(topk or None))
    torch.cuda.synchronize()
    aps, r1s = res[0].cpu().double(), res[1].cpu().double()
    print(f"eval {precision} n={n} d={d}: MAP gpu={aps.mean():.6f} oracle={aps_o.mean():.6f} "
          f"MR1 gpu={r1s.mean():.4f} oracle={r1_o.mean():.4f} max|dAP|={(aps - aps_o).abs().max():.2e} "
          f"n(R1 differ)={(r1s != r1_o).sum().item()} oracle_s={t_or:.1f}", flush=True)
    if topk:
        _, _, tk_idx_o, tk_sim_o = oev.evaluate_argsort(s["c"][:512], s["i"][:512], s["z"][:512], s["c"], s["i"], s["z"], topk=topk)
        idx, sim = res[2][:512].cpu(), res[3][:512].cpu()
        print(f"  topk={topk}: idx mismatch={(idx != tk_idx_o).sum().item()} of {idx.numel()} "
              f"max|dsim|={(sim - tk_sim_o).abs().max():.2e}", flush=True)


def stage_time(precision, n, d, reps=3, topk=0, sigma=2.4):
    s = synth.make_eval_set(n, d, seed=0, device="cuda", md5_ids=False, sigma=sigma)
    plan = we.EvalPlan(s["c"], s["i"], s["c"], s["i"])
    for _ in range(2):
        out = plan.run(s["z"], s["z"], precision=precision, topk=(topk or None))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        out = plan.run(s["z"], s["z"], precision=precision, topk=(topk or None))
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
    m, r1 = we.mean_metrics(out["sums"])
    print(f"time {precision} n={n} d={d} topk={topk} sigma={sigma}: {ms:.2f} ms  {n * n / ms / 1e6:.1f} Gpairs/s  MAP={m:.4f} MR1={r1:.2f} "
          f"pairs={plan.total_pairs} sweep={plan.last_sweep_ms():.2f} ms", flush=True)


def stage_time_chunked(precision, n, s_chunks, d, redux="min", reps=3, ragged=False):
    base = synth.make_eval_set(n, d, seed=0, device="cuda", md5_ids=False)
    g = torch.Generator(device="cuda").manual_seed(1)
    z = (base["z"][:, None, :] + 0.8 * base["z"].norm(dim=1).mean() / d ** 0.5
         * torch.randn(n, s_chunks, d, generator=g, device="cuda")).contiguous()
    plan = we.EvalPlan(base["c"], base["i"], base["c"], base["i"])
    kw = {}
    if ragged:                                  # 1 .. s valid chunks per track
        lens = torch.randint(1, s_chunks + 1, (n,), generator=g, device="cuda").to(torch.int32)
        kw = dict(q_chunks=lens, c_chunks=lens)
    for _ in range(2):
        out = plan.run(z, z, precision=precision, redux=redux, **kw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        out = plan.run(z, z, precision=precision, redux=redux, **kw)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
    m, r1 = we.mean_metrics(out["sums"])
    rows = n * s_chunks
    print(f"time_chunked{' (ragged)' if ragged else ''} {precision} tracks={n} chunks={s_chunks} d={d} redux={redux}: {ms:.2f} ms  "
          f"{rows * rows / ms / 1e6:.1f} G chunk-pairs/s ({n * n / ms / 1e6:.2f} G track-pairs/s)  MAP={m:.4f} MR1={r1:.2f} "
          f"sweep={plan.last_sweep_ms():.2f} ms", flush=True)


def stage_loss(precision):
    import numpy as np
    from wealy_b200 import losses as wl
    from oracle import losses as ol
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "losses.npz"))
    for name in ("f64", "f32", "f32_big", "nopos", "single"):
        z = torch.from_numpy(G[f"{name}_z"]).float()
        lab = torch.from_numpy(G[f"{name}_label"]); idx = torch.from_numpy(G[f"{name}_idx"])
        rows = torch.from_numpy(G[f"{name}_gradrows"])
        for tag, mod, extra, kw in (("ntx", wl.NTXentLoss(0.1, precision=precision), None, {}),
                                    ("ntx05", wl.NTXentLoss(0.5, precision=precision), None, {}),
                                    ("clews", wl.CLEWSLoss(precision=precision), None, {}),
                                    ("clews_step", wl.CLEWSLoss(gamma=6.0, b=0.5, uniformity_weight=0.8, warmup_steps=100, precision=precision), {"global_step": 9}, {}),
                                    ("clews_nf", wl.CLEWSLoss(precision=precision), None, {"numerically_friendly": False})):
            zz = z.cuda().requires_grad_(True)
            lab_c = lab.cuda().clone()
            loss, logd = mod(lab_c, idx.cuda(), zz, extra=extra, **kw)
            loss.backward()
            torch.cuda.synchronize()
            ref_l = float(G[f"{name}_{tag}_loss"]); ref_g = torch.from_numpy(G[f"{name}_{tag}_grad"]).double()
            g = zz.grad.cpu().double()[rows]
            rel_l = abs(float(loss) - ref_l) / max(abs(ref_l), 1e-12)
            rel_g = ((g - ref_g).norm() / ref_g.norm().clamp_min(1e-30)).item()
            extra_s = ""
            if tag in ("ntx", "clews"):
                for k, v in logd.items():
                    key = f"{name}_{tag}_log_{k}"
                    if key in G.files:
                        extra_s += f" {k}:{abs(float(v) - float(G[key])):.1e}"
                extra_s += f" label_ok={bool((lab_c.cpu().numpy() == G[f'{name}_{tag}_label_after']).all())}"
            print(f"loss {precision} {name}/{tag}: loss={float(loss):.6f} ref={ref_l:.6f} rel={rel_l:.2e} grad_relL2={rel_g:.2e}{extra_s}", flush=True)


def stage_losstime(precision, b, d, dtype="bf16"):
    from wealy_b200 import losses as wl
    dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[dtype]
    s = synth.make_loss_batch(b, d, seed=0, dtype=dt, device="cuda")
    for nm, mod in (("ntxent", wl.NTXentLoss(0.1, precision=precision)), ("clews", wl.CLEWSLoss(precision=precision))):
        z = s["z"].clone().requires_grad_(True)
        for _ in range(3):
            loss, _ = mod(s["label"], s["idx"], z); loss.backward()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 10
        e0.record()
        for _ in range(reps):
            loss, _ = mod(s["label"], s["idx"], z)
        e1.record()
        for _ in range(reps):
            loss, _ = mod(s["label"], s["idx"], z); loss.backward()
        e2.record()
        torch.cuda.synchronize()
        print(f"losstime {nm} {precision} {dtype} b={b} d={d}: fwd {e0.elapsed_time(e1) / reps * 1e3:.1f} us  fwd+bwd {e1.elapsed_time(e2) / reps * 1e3:.1f} us loss={float(loss):.5f}", flush=True)


def stage_simtime(precision, n, d):
    x = torch.randn(n, d, device="cuda")
    for _ in range(2):
        out = wt.pairwise_distance_matrix(x, x, mode="cossim", precision=precision)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = wt.pairwise_distance_matrix(x, x, mode="cossim", precision=precision)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"simtime {precision} n={n} d={d}: {ms:.2f} ms {n * n / ms / 1e6:.1f} Gpairs/s  ({2 * n * n * d / ms / 1e9:.0f} TFLOP/s algorithmic)", flush=True)


if __name__ == "__main__":
    st = sys.argv[1]
    a = sys.argv[2:]
    if st == "sim":
        stage_sim(a[0], int(a[1]), int(a[2]), int(a[3]), *(a[4:5]))
    elif st == "eval":
        stage_eval(a[0], int(a[1]), int(a[2]), int(a[3]) if len(a) > 3 else 0)
    elif st == "time":
        stage_time(a[0], int(a[1]), int(a[2]), 3, int(a[3]) if len(a) > 3 else 0, float(a[4]) if len(a) > 4 else 2.4)
    elif st == "time_chunked":
        stage_time_chunked(a[0], int(a[1]), int(a[2]), int(a[3]), *(a[4:5]), ragged=len(a) > 5 and a[5] == "ragged")
    elif st == "loss":
        stage_loss(a[0])
    elif st == "losstime":
        stage_losstime(a[0], int(a[1]), int(a[2]), *(a[3:4]))
    elif st == "simtime":
        stage_simtime(a[0], int(a[1]), int(a[2]))
