timeout 300 python -m pytest tests/test_gpu_eval.py -x -q 2>&1 | tail -3
WEALY_SYM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_eval.py -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 120 python tools/gpu_diag.py time fp16 100000 1024 2>&1 | tail -1
timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
done
timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 3.5 2>&1 | tail -1
WEALY_SYM_PAIR=1 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
