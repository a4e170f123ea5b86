"""Data-parallel loss: global batch = 4096 x 1024 bf16 anchors PER GPU (BASELINE.json configs[3]
per rank, weak scaling), NT-Xent and CLEWS forward+backward over the process group.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_loss_bench.py
Device time (CUDA events), max over ranks; rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200.dist_losses import DistributedNTXentLoss, DistributedCLEWSLoss  # noqa: E402
from wealy_b200.data import synth  # noqa: E402


def main():
    world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nb, d = 4096, 1024
    s = synth.make_loss_batch(nb * world, d, seed=0, dtype=torch.bfloat16, device=dev)
    sl = slice(rank * nb, (rank + 1) * nb)
    lab, idx = s["label"][sl].clone(), s["idx"][sl].clone()
    out = {"config": f"global batch {nb * world} x {d} bf16 = {nb} anchors per GPU x {world} GPUs", "n_gpus": world}
    from wealy_b200 import _native as N
    from wealy_b200.dist_losses import ShardState, _cfg

    def timed(step, reps=10):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    bg = nb * world
    cfgs = {"ntxent": _cfg(kind=N.LOSS_NTXENT, passes=1, temperature=0.1),
            "clews": _cfg(kind=N.LOSS_CLEWS, passes=1, gamma=8.0, b=1.0, eps=1e-8, epsilon=1e-6, uw=0.5, numerically_friendly=1)}
    for name, mod in (("ntxent", DistributedNTXentLoss(0.1)), ("clews", DistributedCLEWSLoss())):
        z = s["z"][sl].clone().requires_grad_(True)

        def step():
            loss, _ = mod(lab, idx, z)
            loss.backward()
            return loss
        os.environ["WEALY_DP_OVERLAP"] = "1"
        ms, loss = timed(step)
        os.environ["WEALY_DP_OVERLAP"] = "0"
        ms_serial, loss_s = timed(step)
        os.environ["WEALY_DP_OVERLAP"] = "auto"
        # the same shard computed from an already gathered global batch: no collective at all (the compute-only floor)
        one = torch.ones((), device=dev)

        def step_nocomm():
            st = ShardState(cfgs[name], s["z"], s["label"], s["idx"], rank * nb, nb)
            st.forward_local()
            o = st.forward_finish()
            st.backward(one)
            return o
        ms_nc, _ = timed(step_nocomm)
        out[name] = {"fwd_bwd_ms": ms, "fwd_bwd_ms_serial_all_gather": ms_serial, "fwd_bwd_ms_no_communication": ms_nc,
                     "ratio_to_no_communication": ms / ms_nc, "loss": float(loss.detach()),
                     "loss_serial": float(loss_s.detach()),
                     "algorithmic_tflops_all_gpus": 8.0 * bg * bg * d / (ms * 1e-3) / 1e12,
                     "anchors_per_s": bg / (ms * 1e-3)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
