#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag3.log
: > $L
run() { echo "### $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python -m pytest tests -m gpu -x -q
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16 100000 1024
WEALY_EPI_WARPS=4 run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024 0 0.5
run python tools/gpu_diag.py time fp16x3 50000 1024 100
run python bench.py --steps 3
run python bench.py --impl reference --steps 1 --warmup 3 --cpu-queries 256
tail -30 $L
