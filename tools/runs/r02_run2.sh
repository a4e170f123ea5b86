#!/bin/bash
# round-2 GPU call 2: symmetric top-k sweep -- targeted tests, then the whole suite, then the bench with the C5 leg
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -k "topk or config5" ) > gpurun_out/r02b_pytest_topk.log 2>&1
echo "pytest-topk rc=$?" >> gpurun_out/r02b_pytest_topk.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
( time timeout 600 python bench.py --legs c5 ) > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err
( WEALY_SYM_TOPK=0 timeout 600 python bench.py --legs c5 --no-cpu ) > gpurun_out/r02b_bench_c5_rect.json 2> gpurun_out/r02b_bench_c5_rect.err
tail -5 gpurun_out/r02b_pytest_topk.log
tail -5 gpurun_out/r02b_pytest.log
