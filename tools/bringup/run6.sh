timeout 600 python -m pytest tests/test_gpu_eval.py -x -q -k "beyond_16" 2>&1 | tail -40
