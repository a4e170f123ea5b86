python tools/gpu_diag.py time_chunked fp16x3 12500 8 1024 min 2>&1 | tail -1
python tools/gpu_diag.py time_chunked fp16x3 12500 8 1024 meanmin 2>&1 | tail -1
python tools/gpu_diag.py time_chunked fp16x3 25000 4 1024 mean 2>&1 | tail -1
python tools/gpu_diag.py time_chunked fp16x3 6250 16 1024 min 2>&1 | tail -1
WEALY_SYM=0 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
