#!/bin/bash
# round-2 GPU call 61: final validation of the shipped library (GPU suite, smoke, default bench)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02v_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^real" gpurun_out/r02v_pytest.log | head -20
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02v_smoke.log 2>&1
tail -5 gpurun_out/r02v_smoke.log | head -2
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02v_bench_default.json 2> gpurun_out/r02v_bench_default.err
tail -3 gpurun_out/r02v_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02v_bench_default.json') if l.startswith('{')][-1])
print('main value %.1f ms %.2f kernel %.2f frac %.3f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']))
e=d['e2e']; print('e2e %.1f (%.2f ms) route %s copy %s pipelined %.2f' % (e['value'], e['ms_per_step'], e['route'], e['copy_then_compute'] and round(e['copy_then_compute']['ms_per_step'],2), e['pipelined']['ms_per_step']))
print('parity', {k:v for k,v in d['parity'].items() if k!='note'}, e.get('abs_dMAP_vs_resident_path'))
for k in ('c1_shs100k','c3_500k','c5_topk100'):
    c=d[k]; p=(c.get('parity') or c.get('parity_all_queries')); print(k, round(c['ms_per_step'],3), p['item_ranks_out_of_band'], p['item_ranks_exact_mismatches'], p.get('topk_idx_mismatches'))
PY
