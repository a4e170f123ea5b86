#!/bin/bash
# round-2 GPU call 39: chunked tests incl. collisions / singletons on the half sweep; OOM-retry build sanity
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_eval_chunked.py tests/test_gpu_eval.py -x -q -k "chunked or collision or scratch or pipeline" ) > gpurun_out/r02s2_pytest.log 2>&1
tail -15 gpurun_out/r02s2_pytest.log | cut -c1-300
