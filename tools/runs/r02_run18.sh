#!/bin/bash
# round-2 GPU call 18: fused redux kernel + full suite
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02q_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02q_pytest.log | head -30
