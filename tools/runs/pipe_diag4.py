"""Diagnostic: kineto trace of EvalPipeline rounds; prints every CPU / CUDA activity longer than 40 ms."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    s = synth.make_eval_set(100_000, 1024, seed=0, device=dev, md5_ids=False)
    c, i, z = s["c"], s["i"], s["z"]
    z_h, c_h, i_h = z.cpu().pin_memory(), c.cpu().pin_memory(), i.cpu().pin_memory()
    del s, z
    for _ in range(3):
        we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h, precision="fp16x3")
    pipe = we.EvalPipeline(precision="fp16x3")
    pipe.result(pipe.submit(c_h, i_h, z_h))
    slow = 0
    for rnd in range(12):
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            torch.cuda.synchronize()
            prev = None
            rows = []
            for _ in range(5):
                t0 = time.perf_counter()
                t = pipe.submit(c_h, i_h, z_h)
                if prev is not None:
                    pipe.result(prev)
                prev = t
                rows.append(round((time.perf_counter() - t0) * 1e3, 1))
            pipe.result(prev)
            torch.cuda.synchronize()
        print("round", rnd, rows, flush=True)
        if max(rows) > 60:
            slow += 1
            evs = sorted(prof.events(), key=lambda e: -max(e.cpu_time_total, e.device_time_total))
            for e in evs[:12]:
                print("   %-60s cpu %.1f ms  device %.1f ms  [%s]" % (e.name[:60], e.cpu_time_total / 1e3, e.device_time_total / 1e3, e.device_type), flush=True)
            if slow >= 2:
                break
    pipe.close()


if __name__ == "__main__":
    main()
