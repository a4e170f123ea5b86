#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag14.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python -m pytest tests/test_gpu_eval.py -q
run python bench.py --steps 5 --no-cpu
run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3
tail -5 $L
