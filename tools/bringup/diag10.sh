#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag10.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python -m pytest tests -m gpu -q
run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_TILES_PER_UNIT=4 run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_TILES_PER_UNIT=16 run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_TILES_PER_UNIT=1000 run python tools/gpu_diag.py time fp16x3 100000 1024
WEALY_GROUP_ROWS=74 run python tools/gpu_diag.py time fp16x3 100000 1024
run python tools/gpu_diag.py time fp16x3 100000 1024 0 0.5
run python tools/gpu_diag.py time fp16 100000 1024
run python tools/gpu_diag.py time fp16x3 50000 1024 100
run python bench.py --steps 3 --no-cpu
tail -30 $L
