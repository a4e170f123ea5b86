"""-m gpu: multi-GPU evaluation over real NCCL (two ranks; skipped on a single-GPU box): the sharded symmetric sweep
with the all-reduce of the rank counters, the sharded upload + NVLink all-gather of host embeddings, and the
query-partitioned general path -- each against the single-GPU result of the same problem."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from wealy_b200 import evaluation as we, dist as wd
        from wealy_b200.data import synth
        s = synth.make_eval_set(5000, 128, seed=17)
        c, i, z = s["c"].to(dev), s["i"].to(dev), s["z"].to(dev)
        ref = we.EvalPlan(c, i, c, i).run(z, z)
        ok = True
        # all-vs-all, device tensors
        out = wd.evaluate_all_vs_all(c, i, z)
        ok &= torch.equal(out["aps"], ref["aps"]) and torch.equal(out["r1s"], ref["r1s"])
        out["plan"].close()
        # all-vs-all, HOST embeddings: every rank uploads half the rows, one all-gather replicates them
        out = wd.evaluate_all_vs_all(s["c"], s["i"], s["z"].pin_memory())
        ok &= torch.equal(out["aps"], ref["aps"]) and torch.equal(out["r1s"], ref["r1s"])
        out["plan"].close()
        # general path: queries partitioned, corpus replicated, results gathered
        q = slice(0, 1001)
        ref_q = we.EvalPlan(c[q], i[q], c, i).run(z[q], z, topk=7)
        got = wd.evaluate_sharded(c[q], i[q], z[q], c, i, z, topk=7)
        ok &= torch.equal(got["aps"], ref_q["aps"]) and torch.equal(got["r1s"], ref_q["r1s"])
        ok &= torch.equal(got["topk_idx"], ref_q["topk_idx"])
        m, r1 = we.mean_metrics(ref_q["sums"])
        ok &= abs(got["map"] - m) < 1e-9 and abs(got["mr1"] - r1) < 1e-6 and got["count"] == 1001
        torch.cuda.synchronize()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_two_ranks_match_one_gpu():
    import torch.multiprocessing as mp
    port = 29900 + (os.getpid() % 90)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
