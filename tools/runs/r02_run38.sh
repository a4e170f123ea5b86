#!/bin/bash
# round-2 GPU call 38: top-k finalize with length-sized exact selection (parity + C5 stage time), pipeline chunked test
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -k "topk or c5 or pipeline or scratch or fallback" ) > gpurun_out/r02r_pytest.log 2>&1
tail -3 gpurun_out/r02r_pytest.log
( timeout 600 python bench.py --legs main,c5 --no-cpu --steps 5 --warmup 3 ) > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02r_bench.json') if l.startswith('{')][-1])
c=d['c5_topk100']
print('c5 %.2f ms sweep %.2f' % (c['ms_per_step'], c['sweep_ms']), {k:round(v['ms'],3) for k,v in c['stages'].items() if isinstance(v,dict)}, c.get('parity'))
PY
tail -2 gpurun_out/r02r_bench.err
