python -m pytest tests -m gpu -x -q 2>&1 | tail -4
WEALY_SYM_LEVELS=3 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_l3.log 2>&1; tail -1 gpurun_out/bench_l3.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
WEALY_SYM_LEVELS=3 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_sym2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
