#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag16.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024 0 0.5
run python tools/gpu_diag.py time fp16 100000 1024
tail -30 $L
