#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag18.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python -m pytest tests/test_gpu_eval.py -q -x
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16 100000 1024
WEALY_SYM=0 run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_SYM=0 run python tools/gpu_diag.py time fp16 100000 1024
run python tools/gpu_diag.py time fp16x3 50000 1024 100
tail -30 $L
