python -m pytest tests/test_gpu_eval.py -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'])"
