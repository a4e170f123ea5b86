#!/usr/bin/env python
"""bench.py -- headline benchmark of the WEALY retrieval-and-scoring hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16x3|fp16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): similarity + MAP evaluation throughput in Gpairs/s, pairs = Nq * Nc per
evaluate.  One *step* = one full fused evaluation (prep -> relevant thresholds -> tcgen05 sweep with
mask / rank-count epilogue -> AP reduce -> MAP) of the workload:

  N = 1 : BASELINE.json configs[1] -- Discogs-VI-YT-test-shaped all-vs-all, 100 000 x 1024-d fp32
          synthetic embeddings, clique sizes bootstrapped from the shipped SHS100K-TEST split.
  N > 1 : the same all-vs-all grown so that every GPU keeps 1e10 pairs (weak scaling):
          N_total = 100 000 * sqrt(N) tracks, corpus replicated; every rank sweeps the row blocks
          rank (mod N) of the symmetric problem and the per-(query, relevant item) rank counters are
          summed with ONE NCCL all-reduce per step (the path's only exchange), after which every
          rank holds the complete AP / R1.

`value`  : whole-job Gpairs/s with embeddings and ids already resident in HBM (id-only plan built
           once, outside the timed region).
`e2e`    : the same metric through the public API `wealy_b200.evaluation.evaluate()` with HOST
           (pinned) buffers: every step pays the host->device copy of embeddings and ids, the plan
           build, the evaluation and the device->host read of per-query AP / R1.
`--impl reference`: the reference's CPU implementation of the path (it is pure Python / torch, so
           the timed code is the oracle port: torch matmul similarity + per-query argsort
           evaluator) on all host cores, on a bounded query sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "similarity+MAP eval Gpairs/s at 1/2/4/8 B200 (roofline frac); MAP parity"
UNIT = "Gpairs/s"
BASE_N = 100_000
DIM = 1024
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        p["_source"] = "measured"
        return p
    p = dict(FALLBACK_PEAKS)
    p["_source"] = "fallback"
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
            }
            self.ok = True
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                self._halt.wait(0.02)
        except Exception:
            self.ok = False

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path (oracle port), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(z, c, i, n_queries, threads=None):
    """Scores the first `n_queries` queries against the full corpus with the reference's CPU
    arithmetic (torch matmul cosine similarity, lib/tensor_ops.py:167-173, + per-query argsort
    evaluator).  -> (seconds, aps, r1s)"""
    import torch
    from oracle import evaluator as oev
    if threads:
        torch.set_num_threads(threads)
    t0 = time.perf_counter()
    aps, r1s = oev.evaluate_argsort(c[:n_queries], i[:n_queries], z[:n_queries], c, i, z)
    return time.perf_counter() - t0, aps, r1s


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from wealy_b200.data import synth
    world = args.gpus
    n_total = int(args.tracks) if args.tracks else int(round(BASE_N * (world ** 0.5)))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s = synth.make_eval_set(n_total, DIM, seed=0, md5_ids=False)
    sample = args.cpu_queries
    times = []
    for it in range(args.warmup + args.steps):
        lo = (it * sample) % max(1, n_total - sample)
        t0 = time.perf_counter()
        from oracle import evaluator as oev
        oev.evaluate_argsort(s["c"][lo:lo + sample], s["i"][lo:lo + sample], s["z"][lo:lo + sample],
                             s["c"], s["i"], s["z"])
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    value = sample * n_total / sec / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"all-vs-all cosine similarity + MAP/MR1, {n_total} x {DIM} fp32 "
                               f"(BASELINE.json configs[1] shape, SHS100K-TEST clique sizes)",
                   "step_sample": f"{sample} queries x {n_total} candidates per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} queries x {n_total} candidates, torch CPU matmul + per-query argsort"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from wealy_b200 import evaluation as we, dist as wd
    from wealy_b200.data import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_total = int(args.tracks) if args.tracks else int(round(BASE_N * (world ** 0.5)))
    lo, hi = wd.shard_range(n_total, rank, world)
    s = synth.make_eval_set(n_total, DIM, seed=0, device=dev, md5_ids=False)
    z, c, i = s["z"], s["c"], s["i"]
    nq = n_total if world == 1 else (hi - lo)          # queries whose row blocks this rank sweeps (about)
    pairs_total = float(n_total) * float(n_total)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- value: inputs resident in HBM, plan built once
    plan = we.EvalPlan(c, i, c, i, device=dev)

    def step_resident():
        if world == 1:
            return plan.run(z, z, precision=args.precision)["sums"]
        plan.sweep_shard(z, rank, world, precision=args.precision)
        dist.all_reduce(plan.counts_tensor())          # int32 rank counters, SUM over ranks
        return plan.finish()["sums"]

    for _ in range(args.warmup):
        step_resident()
    sync_all()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sums = step_resident()
    e1.record()
    sync_all()
    clocks = sampler.finish()
    ms_total = e0.elapsed_time(e1)
    last_sweep_ms = plan.last_sweep_ms()
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = pairs_total / (ms_step * 1e-3) / 1e9
    s_host = sums.double().cpu()
    gpu_map, gpu_mr1 = float(s_host[0] / s_host[2]), float(s_host[1] / s_host[2])

    # ---- e2e: public API with HOST pinned buffers, copies + plan build + result read every step
    z_h, c_h, i_h = z.cpu().pin_memory(), c.cpu().pin_memory(), i.cpu().pin_memory()
    # per rank: its 1/world slice of the embeddings (replicated afterwards by one NVLink all-gather) + all ids
    h2d = -(-n_total // world) * DIM * 4 + c_h.numel() * 8 + i_h.numel() * 8
    aps_h = torch.empty(n_total, dtype=torch.float32).pin_memory()
    r1s_h = torch.empty(n_total, dtype=torch.float32).pin_memory()
    d2h = 2 * n_total * 4

    def step_e2e():
        # public API, host tensors in, host tensors out: upload, id plan, evaluation, read-back every step
        if world == 1:
            aps, r1s = we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h, precision=args.precision)
        else:
            out = wd.evaluate_all_vs_all(c_h, i_h, z_h, precision=args.precision)
            aps, r1s = out["aps"], out["r1s"]
        aps_h.copy_(aps, non_blocking=True)
        r1s_h.copy_(r1s, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        if world > 1:
            out["plan"].close()

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    sync_all()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(e2e_steps):
        step_e2e()
    e3.record()
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(e2.elapsed_time(e3), wall_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e_value = pairs_total / (e2e_ms * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused sweep), measured live with CUDA events on its stream
    peaks = load_peaks()
    passes = 3 if args.precision == "fp16x3" else 1
    # algorithmic work of this rank's launch: its share of the N_total^2 pairs, 2*D flop each
    algo_flops = 2.0 * (float(n_total) * n_total / world) * DIM
    achieved = algo_flops / (last_sweep_ms * 1e-3) / 1e12
    # tensor-core work actually issued: the symmetric sweep contracts only the 128 x 256 tiles that reach above
    # the diagonal (row block rb needs column tiles >= rb / 2), times the passes of the precision mode
    n_rb, n_ct = -(-n_total // 128), -(-n_total // 256)
    tiles = sum(max(0, n_ct - rb // 2) for rb in range(rank, n_rb, world))
    executed = 2.0 * 128 * 256 * DIM * tiles * passes / (last_sweep_ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.precision)
        except Exception:
            traffic = None
    roofline = {
        "kernel": "gemm_kernel<EvalSymEpi> (symmetric tcgen05 similarity sweep over clique-sorted rows + mask + rank-count epilogue)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic,
        "peak_kind": f"{peaks['_source']} dense bf16/fp16 cuBLAS, sustained (kernel timed inside a long step); "
                     f"burst = {peaks['bf16_tflops']}",
        "kernel_ms": last_sweep_ms,
        "executed_tflops": executed,
        "frac_executed": executed / peak,
        "note": "achieved = algorithmic flops (2*D per scored pair, both directions of the symmetric sweep) / "
                "kernel time; executed_tflops = tensor-core work issued: half the tiles (symmetry) x 3 passes in "
                "the fp16x3 parity mode (hi*hi + hi*lo + lo*hi)",
    }

    # ---- CPU baseline (oracle port) on a bounded sample of the same workload, rank 0, N = 1 only
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        nqs = args.cpu_queries
        z_c, c_c, i_c = z.cpu(), c.cpu(), i.cpu()
        sec, aps_o, r1_o = cpu_reference_sample(z_c, c_c, i_c, nqs)
        cpu = {"value": nqs * n_total / sec / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"first {nqs} queries x {n_total} candidates in {sec:.1f} s: torch CPU matmul cosine "
                         f"similarity + per-query argsort AP/R1 (oracle/evaluator.py)"}
        res = plan.run(z, z, precision=args.precision)
        aps_g = res["aps"][:nqs].double().cpu()
        r1_g = res["r1s"][:nqs].double().cpu()
        parity = {"sample_queries": nqs,
                  "abs_dMAP": abs(float(aps_g.mean()) - float(aps_o.mean())),
                  "abs_dMR1": abs(float(r1_g.mean()) - float(r1_o.mean())),
                  "max_abs_dAP": float((aps_g - aps_o).abs().max()),
                  "r1_mismatches": int((r1_g != r1_o).sum())}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 tensor-core hi/lo split x3, f32 accumulate" if passes == 3 else "f16 tensor-core, f32 accumulate",
        "data": "synthetic",
        "config": {
            "workload": f"all-vs-all cosine similarity + self/clique mask + rank + AP/MAP/MR1, {n_total} x {DIM} "
                        f"fp32 embeddings (BASELINE.json configs[1] shape at 1 GPU; N_total = 100000*sqrt(n_gpus), "
                        f"SHS100K-TEST clique-size bootstrap)",
            "queries_per_gpu": nq, "candidates": n_total, "pairs_per_step": pairs_total,
            "parallelism": ("single GPU" if world == 1 else
                            f"row blocks of the symmetric sweep dealt round-robin to {world} ranks, corpus replicated, "
                            f"one NCCL all-reduce of the int32 rank counters per step"),
            "precision": args.precision,
            "l2": "operand planes %.0f MB per step >> 126 MB L2 (no flush needed)" % (n_total * DIM * 2 * (2 if passes == 3 else 1) / 1e6),
            "map": gpu_map, "mr1": gpu_mr1,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms, "api": ("wealy_b200.evaluation.evaluate" if world == 1 else "wealy_b200.dist.evaluate_all_vs_all")
                       + " (host pinned tensors in, host out)"},
        "gpu_launches": 5 * args.steps,   # prep, pos_pairs, pos_sort, fused sweep, ap_reduce per step
        "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks", type=int, default=0,
                    help="override the corpus size (e.g. 500000 = BASELINE configs[2] on 8 GPUs); default: 100000 * sqrt(gpus)")
    ap.add_argument("--precision", default=os.environ.get("WEALY_PRECISION", "fp16x3"), choices=["fp16x3", "fp16"])
    ap.add_argument("--cpu-queries", type=int, default=512, help="queries in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
