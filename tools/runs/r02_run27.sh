#!/bin/bash
# round-2 GPU call 27: tracks K_pos with candidate packing (parity + stage time)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_eval_chunked.py -x -q ) > gpurun_out/r02w_pytest.log 2>&1
tail -3 gpurun_out/r02w_pytest.log
( timeout 600 python bench.py --legs main,f1 --no-cpu --steps 5 --warmup 3 ) > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02w_bench.json') if l.startswith('{')][-1])
f=d['f1_chunked']; print('f1', round(f['ms_per_step'],2), round(f['full_rectangle_ms_per_step'],2), f['rectangle_vs_half_identical'], f['stages_ms'], f['parity'])
print('pipelined', {k:v for k,v in d['e2e']['pipelined'].items() if k!='api'}, 'single', d['e2e']['ms_per_step'])
PY
tail -3 gpurun_out/r02w_bench.err
( timeout 300 python tools/chunked_bench.py ) > gpurun_out/r02w_chunked.json 2>&1; cut -c1-1500 gpurun_out/r02w_chunked.json
