timeout 120 python tools/gpu_diag.py time fp16x3 50000 2048 100 2>&1 | tail -1
timeout 120 python tools/gpu_diag.py time fp16x3 50000 2048 0 2>&1 | tail -1
WEALY_SYM=0 timeout 120 python tools/gpu_diag.py time fp16x3 50000 2048 0 2>&1 | tail -1
timeout 120 python tools/gpu_diag.py time fp16x3 50000 1024 100 2>&1 | tail -1
timeout 120 python tools/gpu_diag.py time fp16x3 50000 1024 10 2>&1 | tail -1
