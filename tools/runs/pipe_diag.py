"""Diagnostic: per-request wall times of EvalPipeline at C2, before and after a CPU-heavy torch phase."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def run(pipe, c_h, i_h, z_h, n, tag):
    torch.cuda.synchronize()
    prev = None
    rows = []
    t_all = time.perf_counter()
    for k in range(n):
        t0 = time.perf_counter()
        t = pipe.submit(c_h, i_h, z_h)
        t1 = time.perf_counter()
        if prev is not None:
            pipe.result(prev)
        t2 = time.perf_counter()
        prev = t
        rows.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
    pipe.result(prev)
    torch.cuda.synchronize()
    print(tag, "total/req %.2f ms" % ((time.perf_counter() - t_all) * 1e3 / n), "(submit, result) ms:", rows, flush=True)


def main():
    dev = torch.device("cuda", 0)
    s = synth.make_eval_set(100_000, 1024, seed=0)
    c_h, i_h, z_h = s["c"].pin_memory(), s["i"].pin_memory(), s["z"].pin_memory()
    for _ in range(3):
        a, r = we.evaluate(c_h, i_h, z_h, c_h, i_h, z_h)
        torch.cuda.synchronize()
    pipe = we.EvalPipeline()
    pipe.result(pipe.submit(c_h, i_h, z_h))
    run(pipe, c_h, i_h, z_h, 6, "fresh   ")
    run(pipe, c_h, i_h, z_h, 6, "again   ")
    # a CPU-heavy phase like the bench's parity / cpu_baseline legs
    x = torch.randn(4096, 4096)
    t0 = time.perf_counter()
    for _ in range(20):
        y = x @ x
        idx = torch.argsort(y, dim=1)
    print("cpu phase %.1f s, threads %d" % (time.perf_counter() - t0, torch.get_num_threads()), flush=True)
    run(pipe, c_h, i_h, z_h, 6, "aftercpu")
    zd = s["z"].double()
    w = zd[:256] @ zd.T
    del zd, w
    run(pipe, c_h, i_h, z_h, 6, "afterf64")
    pipe.close()
    pipe2 = we.EvalPipeline()
    pipe2.result(pipe2.submit(c_h, i_h, z_h))
    run(pipe2, c_h, i_h, z_h, 6, "newpipe ")


if __name__ == "__main__":
    main()
