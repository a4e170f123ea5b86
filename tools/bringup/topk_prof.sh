ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_topk python tools/gpu_diag.py time fp16x3 50000 2048 100 > gpurun_out/ncu_topk.log 2>&1
echo rc=$?
