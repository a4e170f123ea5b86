timeout 60 python tools/gpu_diag.py eval fp16x3 3000 256 2>&1 | tail -3
timeout 60 python tools/gpu_diag.py eval fp16x3 10547 1024 2>&1 | tail -2
timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -2
WEALY_SYM_PAIR=0 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
