#!/bin/bash
# round-2 GPU call 31: big-block free list (scratch test, pipeline stalls, bench e2e)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -k "scratch or pipeline or topk or fallback or c5 or kat" ) > gpurun_out/r02y_pytest.log 2>&1
tail -4 gpurun_out/r02y_pytest.log
timeout 300 python tools/pipe_diag3.py --close 2>&1 | tail -6
timeout 300 python tools/pipe_diag3.py 2>&1 | tail -4
for k in 1 2 3; do
( timeout 600 python bench.py --legs main --no-cpu --steps 5 --warmup 3 ) > gpurun_out/r02y_bench$k.json 2> gpurun_out/r02y_bench$k.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02y_bench$k.json') if l.startswith('{')][-1])
print('value %.1f pipelined' % d['value'], {k:v for k,v in d['e2e']['pipelined'].items() if k!='api'}, 'single', d['e2e']['ms_per_step'])
PY
done
