#!/bin/bash
# round-2 GPU call 52: validation of the re-built library after the scheduling-group change (GPU suite with durations, smoke, default bench)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/r02y_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^real" gpurun_out/r02y_pytest.log | head -20
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02y_smoke.log 2>&1
tail -5 gpurun_out/r02y_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err
tail -4 gpurun_out/r02y_bench_default.err
cut -c1-400 gpurun_out/r02y_bench_default.json
