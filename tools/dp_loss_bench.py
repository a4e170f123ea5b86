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
    for name, mod in (("ntxent", DistributedNTXentLoss(0.1)), ("clews", DistributedCLEWSLoss())):
        z = s["z"][sl].clone().requires_grad_(True)
        for _ in range(3):
            loss, _ = mod(lab, idx, z)
            loss.backward()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            loss, _ = mod(lab, idx, z)
            loss.backward()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        bg = nb * world
        out[name] = {"fwd_bwd_ms": ms, "loss": float(loss.detach()),
                     "algorithmic_tflops_all_gpus": 8.0 * bg * bg * d / (ms * 1e-3) / 1e12,
                     "anchors_per_s": bg / (ms * 1e-3)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
