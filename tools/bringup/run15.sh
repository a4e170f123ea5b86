python -m pytest tests/test_gpu_eval.py -x -q 2>&1 | tail -3
for i in 1 2; do python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1; done
python tools/gpu_diag.py time fp16 100000 1024 2>&1 | tail -1
WEALY_TILES_PER_UNIT=16 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
WEALY_TILES_PER_UNIT=4 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
