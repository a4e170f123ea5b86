CMD="python bench.py --steps 2 --warmup 3 --no-cpu --precision fp16"
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_fp16_b $CMD > gpurun_out/ncu_full2.log 2>&1
echo rc=$?
