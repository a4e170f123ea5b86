#!/bin/bash
# ncu evidence for the bench command (run under gpurun; results copied to profiles/ afterwards)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -o gpurun_out/prof_fp16x3 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu --precision fp16"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -o gpurun_out/prof_fp16 $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "full2 rc=$?"
tail -2 gpurun_out/plain.log gpurun_out/plain3.log
ls -la gpurun_out | tail -12
