#!/bin/bash
# round-2 GPU call 5: group-max top-k pre-pass, pair kernel with relaxed hand-over arrivals
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02e_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02e_pytest.log | head -20
B="python bench.py --legs main --no-cpu --steps 20 --warmup 5"
run() { name=$1; shift; ( env "$@" timeout 300 $B ) > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02e_$name.json') if l.startswith('{')][-1])
    print('$name', 'value %.1f ms %.2f kernel %.2f e2e %.1f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['run']['map'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/r02e_$name.err').read()[-600:])
PY
}
run single WEALY_SYM_PAIR=0
run pair_static12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=0
run pair_static8 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=0 WEALY_PAIR_EPI_WARPS=8
run pair_dyn12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1
run pair_dyn8 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1 WEALY_PAIR_EPI_WARPS=8
run pair_dyn12_t4 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1 WEALY_TILES_PER_UNIT=4
run pair_dyn12_t16 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1 WEALY_TILES_PER_UNIT=16
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --sigma 4.0"
run hard_pair_dyn12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1
run hard_pair_static12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=0
( timeout 600 python bench.py --legs c5 --no-cpu ) > gpurun_out/r02e_bench_c5.json 2> gpurun_out/r02e_bench_c5.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02e_bench_c5.json') if l.startswith('{')][-1]); c=d['c5_topk100']
print('c5', c['ms_per_step'], c['sweep_ms'], c['topk_path'], {k:v['ms'] for k,v in c['stages'].items() if isinstance(v,dict)})
PY
