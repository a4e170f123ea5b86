"""Diagnostic: where the first requests of an EvalPipeline spend their time (phases of submit(), synchronised)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def traced_submit(self, c, i, z):
    T = [time.perf_counter()]
    def mark(sync=True):
        if sync:
            torch.cuda.synchronize()
        T.append(time.perf_counter())
    z, c, i = torch.as_tensor(z), torch.as_tensor(c).long(), torch.as_tensor(i).long()
    slot = self.slots[self.count % self.depth]
    if slot["plan"] is not None:
        slot["done"].synchronize()
        slot["plan"].close()
        slot["plan"] = None
    mark()                                   # 1 close
    self._buffers(slot, z, c, i)
    mark()                                   # 2 buffers
    compute = torch.cuda.current_stream(self.device)
    with torch.cuda.stream(self.copy_stream):
        slot["z"].copy_(z, non_blocking=True)
        slot["c"].copy_(c, non_blocking=True)
        slot["i"].copy_(i, non_blocking=True)
        uploaded = torch.cuda.Event()
        uploaded.record(self.copy_stream)
    mark(False)                              # 3 copies issued
    mark()                                   # 4 copies done
    compute.wait_event(uploaded)
    plan = we.EvalPlan(slot["c"], slot["i"], slot["c"], slot["i"], device=self.device)
    mark()                                   # 5 plan
    res = plan.run(slot["z"], slot["z"], eps=self.eps, precision=self.precision)
    mark()                                   # 6 run
    slot["aps"].copy_(res["aps"], non_blocking=True)
    slot["r1s"].copy_(res["r1s"], non_blocking=True)
    slot["done"] = torch.cuda.Event()
    slot["done"].record(compute)
    mark()                                   # 7 d2h
    slot["plan"], slot["sums"] = plan, res["sums"]
    self.count += 1
    names = ["close", "buffers", "issue", "h2d", "plan", "run", "d2h"]
    print("   ", {n: round((T[k + 1] - T[k]) * 1e3, 2) for k, n in enumerate(names)}, flush=True)
    return self.count - 1


def main():
    s = synth.make_eval_set(100_000, 1024, seed=0)
    c_h, i_h, z_h = s["c"].pin_memory(), s["i"].pin_memory(), s["z"].pin_memory()
    for _ in range(3):
        we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h)
        torch.cuda.synchronize()
    we.EvalPipeline.submit = traced_submit
    pipe = we.EvalPipeline()
    print("fresh pipeline", flush=True)
    for _ in range(5):
        pipe.submit(c_h, i_h, z_h)
    x = torch.randn(4096, 4096)
    for _ in range(20):
        y = x @ x
        idx = torch.argsort(y, dim=1)
    print("after a CPU phase", flush=True)
    for _ in range(4):
        pipe.submit(c_h, i_h, z_h)
    time.sleep(3)
    print("after sleeping 3 s", flush=True)
    for _ in range(3):
        pipe.submit(c_h, i_h, z_h)
    print("evaluate() after sleeping 3 s", flush=True)
    time.sleep(3)
    for _ in range(3):
        t0 = time.perf_counter()
        we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h)
        torch.cuda.synchronize()
        print("    %.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)


if __name__ == "__main__":
    main()
