#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag19.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python tools/gpu_diag.py time fp16x3 50000 1024 100
run python tools/gpu_diag.py time fp16x3 50000 1024 100
run python tools/gpu_diag.py time fp16x3 50000 2048 100
run python tools/gpu_diag.py time fp16x3 100000 1024
tail -30 $L
