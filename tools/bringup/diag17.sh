#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag17.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
WEALY_EVAL_EPI_WARPS=16 run python -m pytest tests/test_gpu_eval.py -q -x
WEALY_EVAL_EPI_WARPS=16 run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_EVAL_EPI_WARPS=16 run python tools/gpu_diag.py time fp16x3 100000 1024
tail -30 $L
