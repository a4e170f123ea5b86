"""Config 4 (BASELINE.json configs[3]): contrastive loss forward / backward, batch 4096 x 1024 bf16.
Times the fused CUDA modules (CUDA events, L2 flushed between iterations) and, on the host cores,
the oracle port of the reference modules.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402
from wealy_b200 import losses as wl  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def gpu_time(mod, s, reps, backward, flush):
    z = s["z"].clone().requires_grad_(True)
    for _ in range(3):
        loss, _ = mod(s["label"], s["idx"], z)
        if backward:
            loss.backward()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()                      # > 126 MB: evicts L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, _ = mod(s["label"], s["idx"], z)
        if backward:
            loss.backward()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], float(loss.detach())


def graph_time(mod, s, reps, flush):
    """The same forward + backward captured once in a CUDA graph and replayed (the step is launch bound: ~10 short
    kernels): device time per replay."""
    z = s["z"].clone().requires_grad_(True)
    lab, idx = s["label"].clone(), s["idx"].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            z.grad = None
            loss, _ = mod(lab, idx, z)
            loss.backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    z.grad = None
    with torch.cuda.graph(g):
        loss, _ = mod(lab, idx, z)
        loss.backward()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], float(loss.detach()), z.grad.detach().clone()


def main():
    b, d = 4096, 1024
    cpu = "--no-cpu" not in sys.argv
    precision = None          # auto: fp16x3 for fp32 inputs, one fp16 pass for fp16 / bf16 inputs
    for a in sys.argv[1:]:
        if a.startswith("--precision="):
            precision = a.split("=")[1]
    s = synth.make_loss_batch(b, d, seed=0, dtype=torch.bfloat16, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {"config": f"loss fwd/bwd, batch {b} x {d} bf16, 4 items per clique", "precision": precision or "auto (fp16 for bf16 input)"}
    for name, mod in (("ntxent", wl.NTXentLoss(0.1, precision=precision)), ("clews", wl.CLEWSLoss(precision=precision))):
        f, lv = gpu_time(mod, s, 20, False, flush)
        fb, _ = gpu_time(mod, s, 20, True, flush)
        passes = 3 if precision == "fp16x3" else 1  # auto resolves to 1 pass for the bf16 batch
        try:
            gms, glv, ggrad = graph_time(mod, s, 20, flush)
            z2 = s["z"].clone().requires_grad_(True)
            l2, _ = mod(s["label"].clone(), s["idx"], z2)
            l2.backward()
            graph = {"fwd_bwd_graph_ms": gms, "graph_loss": glv,
                     "graph_grad_rel_diff": float((ggrad.float() - z2.grad.float()).norm() / z2.grad.float().norm())}
        except Exception as e:  # report, do not hide
            graph = {"fwd_bwd_graph_ms": None, "graph_error": repr(e)[:200]}
        out[name] = {"fwd_ms": f, "fwd_bwd_ms": fb, "loss": lv, **graph,
                     "algorithmic_tflops_fwd_bwd": 8.0 * b * b * d / (fb * 1e-3) / 1e12,
                     "executed_tflops_fwd_bwd": 8.0 * b * b * d * passes / (fb * 1e-3) / 1e12}
    if cpu:
        from oracle import losses as ol
        torch.set_num_threads(os.cpu_count() or 1)
        zc = s["z"].float().cpu()
        lab, idx = s["label"].cpu(), s["idx"].cpu()
        for name, fn in (("ntxent", lambda z: ol.ntxent(lab.clone(), idx, z, 0.1)[0]),
                         ("clews", lambda z: ol.clews(lab.clone(), idx, z)[0])):
            best_f = best_fb = 1e9
            for _ in range(3):
                z = zc.clone().requires_grad_(True)
                t0 = time.perf_counter(); loss = fn(z); t1 = time.perf_counter(); loss.backward(); t2 = time.perf_counter()
                best_f, best_fb = min(best_f, t1 - t0), min(best_fb, t2 - t0)
            out[name]["cpu_port_fwd_ms"] = best_f * 1e3
            out[name]["cpu_port_fwd_bwd_ms"] = best_fb * 1e3
            out[name]["cpu_cores"] = torch.get_num_threads()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
