"""-m gpu: multi-rank evaluation, each case against the single-GPU result of the same problem:
* over real NCCL on two GPUs (skipped on a single-GPU box): the sharded symmetric sweep with the all-reduce of the rank
  counters, the sharded upload + NVLink all-gather of host embeddings, the query-partitioned general path;
* two PROCESSES sharing cuda:0 over gloo (runs on a single-GPU box too): the same sharded sweep / all-reduce /
  finish and the query-partitioned path through real collectives, plus the error agreement across ranks."""
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


def _gloo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(0)                                  # both ranks on the one GPU of the box
    dev = torch.device("cuda", 0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wealy_b200 import evaluation as we, dist as wd
        from wealy_b200.data import synth
        from oracle import evaluator as oev
        s = synth.make_eval_set(3000, 96, seed=19)
        c, i, z = s["c"].to(dev), s["i"].to(dev), s["z"].to(dev)
        ref = we.EvalPlan(c, i, c, i).run(z, z)
        ok = True
        out = wd.evaluate_all_vs_all(c, i, z)                 # sweep_shard -> all_reduce(counters) -> finish
        ok &= torch.equal(out["aps"], ref["aps"]) and torch.equal(out["r1s"], ref["r1s"]) and out["count"] == 3000
        # the summed counters give every relevant item's rank: bands of the oracle on a query sample
        off_g, ranks_g, _ = (t.cpu() for t in out["plan"].ranks())
        qs = torch.arange(rank, 3000, 37)
        off_o, _, exact, lo, hi = oev.rank_bands(s["c"][qs], s["i"][qs], s["z"][qs], s["c"], s["i"], s["z"])
        got = torch.cat([ranks_g[int(off_g[q]):int(off_g[q + 1])] for q in qs.tolist()]).long()
        ok &= bool(((got >= lo) & (got <= hi)).all()) and torch.equal(got[lo == hi], exact[lo == hi])
        out["plan"].close()
        q = slice(0, 701)
        ref_q = we.EvalPlan(c[q], i[q], c, i).run(z[q], z, topk=5)
        got = wd.evaluate_sharded(c[q], i[q], z[q], c, i, z, topk=5)
        ok &= torch.equal(got["aps"], ref_q["aps"]) and torch.equal(got["topk_idx"], ref_q["topk_idx"]) and got["count"] == 701
        # all-vs-all with top-k on several ranks: query-partitioned underneath, lists gathered
        ref_k = we.EvalPlan(c, i, c, i).run(z, z, topk=6)
        got_k = wd.evaluate_all_vs_all(c, i, z, topk=6)
        ok &= got_k["topk_idx"].shape == (3000, 6) and float((got_k["topk_sim"] - ref_k["topk_sim"]).abs().max()) <= 2e-6
        ok &= abs(got_k["map"] - float(ref["aps"].double().mean())) <= 1e-6
        # a query without relevant candidates in ONE rank's slice: every rank raises (no rank is left in a collective)
        c_bad = c.clone()
        c_bad[0] = 10_000_000                                 # query 0 (rank 0's slice) loses its clique
        try:
            wd.evaluate_sharded(c_bad[q], i[q], z[q], c, i, z)
            ok = False
        except ValueError:
            pass
        try:
            wd.evaluate_all_vs_all(c_bad, i, z)
            ok = False
        except ValueError:
            pass
        # more ranks than queries: the empty shard contributes nothing
        tiny = wd.evaluate_sharded(c[:1], i[:1], z[:1], c, i, z)
        ref_1 = we.EvalPlan(c[:1], i[:1], c, i).run(z[:1], z)
        ok &= tiny["count"] == 1 and torch.equal(tiny["aps"], ref_1["aps"]) and torch.equal(tiny["r1s"], ref_1["r1s"])
        torch.cuda.synchronize()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_processes_one_gpu_gloo():
    import torch.multiprocessing as mp
    port = 29300 + (os.getpid() % 90)
    ret = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
