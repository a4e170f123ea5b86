#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
( time timeout 600 python bench.py --legs c5,c4 ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02c_pytest.log | head -30
