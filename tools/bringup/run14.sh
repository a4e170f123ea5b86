python -m pytest tests/test_gpu_eval_chunked.py -x -q 2>&1 | tail -3
python tools/gpu_diag.py time_chunked fp16x3 12500 8 1024 min 2>&1 | tail -1
python tools/gpu_diag.py time_chunked fp16x3 25000 4 1024 mean 2>&1 | tail -1
python tools/gpu_diag.py time_chunked fp16x3 6250 16 1024 meanmin 2>&1 | tail -1
python tools/gpu_diag.py time_chunked fp16x3 50000 2 1024 minmean 2>&1 | tail -1
