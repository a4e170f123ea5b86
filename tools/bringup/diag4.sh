#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag4.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python -m pytest tests -m gpu -q
run python tools/gpu_diag.py time fp16x3 50000 1024 100
run python bench.py --steps 3
tail -30 $L
