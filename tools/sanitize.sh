#!/bin/bash
# memcheck on small instances of every kernel family (one compute-sanitizer tool per gpurun call)
mkdir -p gpurun_out
L=gpurun_out/sanitize.log
: > $L
run() { echo "### $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run compute-sanitizer --tool memcheck --error-exitcode 99 python tools/gpu_diag.py eval fp16x3 700 96 10
run compute-sanitizer --tool memcheck --error-exitcode 99 python tools/gpu_diag.py eval fp16 1300 64 100
run compute-sanitizer --tool memcheck --error-exitcode 99 python tools/gpu_diag.py sim fp16x3 130 257 200 fro
run compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_losses.py -q -k "ragged or upstream"
run compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_masked.py -q -k "masked_reductions or long_rows"
grep -E "ERROR SUMMARY|exit=|Invalid|Error" $L | head -40
