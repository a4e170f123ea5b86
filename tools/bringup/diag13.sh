#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/diag13.log
: > $L
run() { echo "### $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python -m pytest tests -m gpu -q
run python bench.py --steps 5
run python tools/gpu_diag.py time fp16x3 500000 1024
run python -c "import __graft_entry__ as g; g.smoke()"
tail -30 $L
