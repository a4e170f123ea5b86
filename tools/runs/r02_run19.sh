#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_masked.py -q ) > gpurun_out/r02r_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02r_pytest.log | head -30
grep -E "^E  " gpurun_out/r02r_pytest.log | sort | uniq -c | sort -rn | head
