#!/bin/bash
# round-2 GPU call 1: full GPU test suite, default bench (all legs), pair-kernel headline, reference arm sanity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench rc=$?" >> gpurun_out/r02_bench_default.err
( WEALY_SYM_PAIR=1 timeout 300 python bench.py --legs main --no-cpu --steps 20 --warmup 5 ) > gpurun_out/r02_bench_pair.json 2> gpurun_out/r02_bench_pair.err
( timeout 300 python bench.py --legs main --no-cpu --steps 20 --warmup 5 ) > gpurun_out/r02_bench_single.json 2> gpurun_out/r02_bench_single.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 3 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -3 gpurun_out/r02_pytest.log
cut -c1-600 gpurun_out/r02_bench_default.json
tail -2 gpurun_out/r02_bench_default.err
