#!/bin/bash
# round-2 GPU call 13: rectangle sweeps on the CTA-pair core (full suite + timings vs the single-CTA core)
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02l_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02l_pytest.log | head -20
( WEALY_RECT_PAIR=1 timeout 600 python tools/rect_bench.py ) > gpurun_out/r02l_rect_pair.json 2> gpurun_out/r02l_rect_pair.err
tail -1 gpurun_out/r02l_rect_pair.json; tail -2 gpurun_out/r02l_rect_pair.err
( WEALY_RECT_PAIR=0 timeout 600 python tools/rect_bench.py ) > gpurun_out/r02l_rect_single.json 2> gpurun_out/r02l_rect_single.err
tail -1 gpurun_out/r02l_rect_single.json; tail -2 gpurun_out/r02l_rect_single.err
