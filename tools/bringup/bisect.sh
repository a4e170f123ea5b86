#!/bin/bash
mkdir -p gpurun_out
L=$PWD/gpurun_out/bisect.log
: > $L
for c in c0bae81 fa748f5 258750a; do
  echo "### $c" >> $L
  (cd bisect/$c && timeout 600 python tools/gpu_diag.py time fp16x3 50000 1024 100 >> $L 2>&1; timeout 600 python tools/gpu_diag.py time fp16x3 50000 1024 100 >> $L 2>&1)
done
echo "### HEAD" >> $L
timeout 600 python tools/gpu_diag.py time fp16x3 50000 1024 100 >> $L 2>&1
cat $L | grep -E "^###|^time"
