#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag12.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
WEALY_SYM=1 run python -m pytest tests/test_gpu_eval.py -q -x
WEALY_SYM=1 run python tools/gpu_diag.py eval fp16x3 3000 1024
WEALY_SYM=1 run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_SYM=1 run python tools/gpu_diag.py time fp16 100000 1024
WEALY_SYM=1 run python tools/gpu_diag.py time fp16x3 100000 1024 0 0.5
WEALY_SYM=1 run python tools/gpu_diag.py time fp16x3 20000 1024
tail -30 $L
