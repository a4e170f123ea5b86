#!/bin/bash
# one ncu --set full capture (with source correlation) of the symmetric fused sweep of the bench command
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_sym $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -1 gpurun_out/plain.log | cut -c1-400
