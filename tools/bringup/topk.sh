python -m pytest tests/test_gpu_eval.py tests/test_gpu_eval_chunked.py -x -q -k "topk" 2>&1 | tail -3
python tools/gpu_diag.py time fp16x3 50000 2048 100 2>&1 | tail -1
python tools/gpu_diag.py time fp16x3 50000 1024 100 2>&1 | tail -1
python tools/gpu_diag.py time fp16x3 50000 1024 10 2>&1 | tail -1
python tools/gpu_diag.py time fp16x3 50000 1024 500 2>&1 | tail -1
