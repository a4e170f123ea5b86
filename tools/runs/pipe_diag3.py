"""Diagnostic: EvalPipeline per-request wall times in the exact setting of bench.py (resident plan alive, evaluate() steps first)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    s = synth.make_eval_set(100_000, 1024, seed=0, device=dev, md5_ids=False)
    c, i, z = s["c"], s["i"], s["z"]
    plan = we.EvalPlan(c, i, c, i, device=dev)
    for _ in range(8):
        res = plan.run(z, z, precision="fp16x3")
    torch.cuda.synchronize()
    z_h, c_h, i_h = z.cpu().pin_memory(), c.cpu().pin_memory(), i.cpu().pin_memory()
    keep_plan = "--close" not in sys.argv
    if not keep_plan:
        plan.close()
    for _ in range(7):
        t0 = time.perf_counter()
        a, r = we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h, precision="fp16x3")
        torch.cuda.current_stream(dev).synchronize()
        print("evaluate %.2f ms" % ((time.perf_counter() - t0) * 1e3), {k: v >> 20 for k, v in we.pool_stats().items()}, flush=True)
    pipe = we.EvalPipeline(precision="fp16x3")
    pipe.result(pipe.submit(c_h, i_h, z_h))
    for rnd in range(4):
        torch.cuda.synchronize()
        prev = None
        rows = []
        t_all = time.perf_counter()
        for _ in range(5):
            t0 = time.perf_counter()
            t = pipe.submit(c_h, i_h, z_h)
            if prev is not None:
                pipe.result(prev)
            prev = t
            rows.append(round((time.perf_counter() - t0) * 1e3, 1))
        pipe.result(prev)
        torch.cuda.synchronize()
        print("round", rnd, "per request %.2f ms" % ((time.perf_counter() - t_all) * 1e3 / 5), rows,
              {k: v >> 20 for k, v in we.pool_stats().items()}, flush=True)
    pipe.close()


if __name__ == "__main__":
    main()
