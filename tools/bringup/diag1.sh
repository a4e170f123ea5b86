#!/bin/bash
# first GPU bring-up: core contraction, then the fused evaluation
mkdir -p gpurun_out
L=gpurun_out/diag1.log
: > $L
run() { echo "### $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L 2>&1
run python tools/gpu_diag.py sim fp16x3 130 257 200
run python tools/gpu_diag.py sim fp16 130 257 200
WEALY_BLOCK_K=32 run python tools/gpu_diag.py sim fp16x3 130 257 200
run python tools/gpu_diag.py sim fp16x3 1000 3000 1024
run python tools/gpu_diag.py sim fp16 1000 3000 1024
run python tools/gpu_diag.py sim fp16x3 300 700 96 dot
run python tools/gpu_diag.py sim fp16x3 300 700 96 fro
run python tools/gpu_diag.py eval fp16x3 2000 128
run python tools/gpu_diag.py eval fp16 2000 128
run python tools/gpu_diag.py eval fp16x3 3000 1024 10
run python tools/gpu_diag.py time fp16x3 20000 1024
run python tools/gpu_diag.py time fp16 20000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16 100000 1024
WEALY_BLOCK_K=32 run python tools/gpu_diag.py time fp16x3 100000 1024
tail -5 $L
