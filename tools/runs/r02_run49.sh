#!/bin/bash
# round-2 GPU call 49: new default scheduling group of the plain pair sweep: C2 / C1 / C3 / sharded emulation
mkdir -p gpurun_out
( timeout 600 python bench.py --legs main,c1,c3 --no-cpu --steps 10 --warmup 3 ) > gpurun_out/r02grp.json 2> gpurun_out/r02grp.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02grp.json') if l.startswith('{')][-1])
print('value %.1f ms %.2f kernel %.2f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))
print('c1', d['c1_shs100k']['ms_per_step'], 'c3', d['c3_500k']['ms_per_step'], d['c3_500k']['gpairs_per_s'])
PY
( WEALY_GROUP_ROWS=37 timeout 600 python bench.py --legs main,c1,c3 --no-cpu --steps 10 --warmup 3 ) > gpurun_out/r02grp_old.json 2> gpurun_out/r02grp_old.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02grp_old.json') if l.startswith('{')][-1])
print('old grouping: value %.1f ms %.2f kernel %.2f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))
print('c1', d['c1_shs100k']['ms_per_step'], 'c3', d['c3_500k']['ms_per_step'], d['c3_500k']['gpairs_per_s'])
PY
( timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_dist_eval.py -x -q ) 2>&1 | tail -2
