set -x
timeout 120 python tools/gpu_diag.py eval fp16x3 3000 256 2>&1 | tail -3
timeout 120 python tools/gpu_diag.py eval fp16x3 10547 1024 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_eval.py -x -q 2>&1 | tail -8
for lv in 2 3 4; do WEALY_SYM_LEVELS=$lv timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1; done
WEALY_SYM_LEVELS=4 timeout 120 python tools/gpu_diag.py time fp16 100000 1024 2>&1 | tail -1
