"""End to end at BASELINE configs[1] through the public API with pinned host tensors: the upload / prep / sweep pipeline
(wealy_eval_run_host) under each part schedule against copy-then-compute (WEALY_HOST_STREAM=0).  One JSON line."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=100000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--parts", default="0,1,2,5,7")
    ap.add_argument("--only-host", action="store_true", help="skip the copy-then-compute runs (for a profiler to attach to)")
    ap.add_argument("--sweep", default="", help='";"-separated settings, each "VAR=value&VAR=value" (WEALY_HOST_* knobs), timed in turn')
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    s = synth.make_eval_set(args.tracks, args.dim, seed=0, device=dev, md5_ids=False)
    z_h, c_h, i_h = s["z"].cpu().pin_memory(), s["c"].cpu().pin_memory(), s["i"].cpu().pin_memory()
    aps_h = torch.empty(args.tracks, dtype=torch.float32).pin_memory()
    r1s_h = torch.empty(args.tracks, dtype=torch.float32).pin_memory()

    def step():
        aps, r1s = we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h)
        aps_h.copy_(aps, non_blocking=True)
        r1s_h.copy_(r1s, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    out = {"tracks": args.tracks, "dim": args.dim, "steps": args.steps, "runs": []}
    # the id plan alone (host ids in, two radix sorts, CSR, host read-backs): exposed in front of the pipeline
    for _ in range(2):
        we.EvalPlan(c_h, i_h, c_h, i_h).close()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        we.EvalPlan(c_h, i_h, c_h, i_h).close()
    torch.cuda.synchronize()
    out["plan_build_ms"] = (time.perf_counter() - t0) * 1e3 / 5
    ref = None
    modes = [f"parts{p}" for p in args.parts.split(",")]
    for mode in modes if args.only_host else ["copy"] + modes + ["copy"]:
        if mode == "copy":
            os.environ["WEALY_HOST_STREAM"] = "0"
        else:
            os.environ["WEALY_HOST_STREAM"] = "1"
            os.environ["WEALY_HOST_PARTS"] = mode[5:]
        step()
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        a = aps_h.clone()
        if ref is None:
            ref = a
        out["runs"].append({"mode": mode, "route": we.last_path(), "ms_per_step": ms,
                            "gpairs_per_s": args.tracks * args.tracks / ms / 1e6,
                            "identical_to_copy_path": bool(torch.equal(a, ref)), "map": float(a.double().mean())})
    if args.sweep:
        os.environ["WEALY_HOST_STREAM"] = "1"
        os.environ.pop("WEALY_HOST_PARTS", None)
        out["sweep"] = []
        for rep in range(2):
            for setting in args.sweep.split(";"):
                kv = dict(t.split("=", 1) for t in setting.split("&") if t)
                for k_, v_ in kv.items():
                    os.environ[k_] = v_
                step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    step()
                ms = (time.perf_counter() - t0) * 1e3 / args.steps
                out["sweep"].append({"setting": setting, "rep": rep, "ms_per_step": ms, "identical": bool(torch.equal(aps_h, ref))})
                for k_ in kv:
                    os.environ.pop(k_, None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
