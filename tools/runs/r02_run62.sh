#!/bin/bash
# round-2 GPU call 62: the re-built library (comment-only change) on the evaluation tests that exercise the host pipeline + smoke
mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_gpu_eval_host.py tests/test_gpu_eval.py -x -q ) > gpurun_out/r02u_pytest.log 2>&1
tail -2 gpurun_out/r02u_pytest.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
