#!/bin/bash
# round-2 GPU call 22: symmetric chunked sweep (parity + timing), EvalPipeline test
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_eval_chunked.py tests/test_gpu_eval.py -x -q -k "chunked or ragged or pipeline or symmetric_sweep_ragged" ) > gpurun_out/r02t_pytest.log 2>&1
tail -5 gpurun_out/r02t_pytest.log
( timeout 600 python tools/rect_bench.py ) > gpurun_out/r02t_rect.json 2> gpurun_out/r02t_rect.err
cat gpurun_out/r02t_rect.json; tail -3 gpurun_out/r02t_rect.err
