#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag2.log
: > $L
run() { echo "### $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python tools/gpu_diag.py eval fp16x3 2000 128
run python tools/gpu_diag.py eval fp16x3 3000 1024 10
run python tools/gpu_diag.py eval fp16 3000 256 100
run python tools/gpu_diag.py loss fp16x3
run python tools/gpu_diag.py loss fp16
run python tools/gpu_diag.py simtime fp16x3 16384 1024
run python tools/gpu_diag.py simtime fp16 16384 1024
run python tools/gpu_diag.py time fp16x3 100000 1024 0 0.5
run python tools/gpu_diag.py time fp16 100000 1024 0 0.5
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16 100000 1024
WEALY_BLOCK_K=32 run python tools/gpu_diag.py time fp16x3 100000 1024 0 0.5
run python tools/gpu_diag.py time fp16x3 50000 1024 100
run python tools/gpu_diag.py losstime fp16x3 4096 1024 bf16
run python tools/gpu_diag.py losstime fp16 4096 1024 bf16
tail -5 $L
