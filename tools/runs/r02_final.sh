#!/bin/bash
# final validation of the round: what the driver runs (GPU suite, smoke, bench, reference arm), outputs kept for profiles/
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02z_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02z_pytest.log | head -20
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02z_smoke.log 2>&1
tail -2 gpurun_out/r02z_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02z_bench_default.json') if l.startswith('{')][-1])
print('main value %.1f ms %.2f kernel %.2f frac %.3f e2e %.1f (%.2f ms) clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks']))
print('traffic', d['roofline']['traffic'], d['roofline']['traffic_note'][:80])
print('stages', {k:(round(v['ms'],3), v['gbs'] and round(v['gbs'])) for k,v in d['roofline']['stages'].items() if isinstance(v,dict)})
print('parity', {k:v for k,v in d['parity'].items() if k!='note'})
for k in ('c1_shs100k','c3_500k','c5_topk100'):
    c=d[k]; p=(c.get('parity') or c.get('parity_all_queries')); print(k, round(c['ms_per_step'],3), c.get('gpairs_per_s'), c.get('roofline_frac'), c.get('topk_path'), p['item_ranks_out_of_band'], p['item_ranks_exact_mismatches'], p.get('topk_idx_mismatches'))
print('c4', {k:(round(v['fwd_ms'],4),round(v['fwd_bwd_eager_ms'],4),round(v['fwd_bwd_graph_ms'],4),v['loss_rel_err_vs_cpu_fp32']) for k,v in d['c4_loss'].items() if isinstance(v,dict)})
f=d['f1_chunked']; print('f1', round(f['ms_per_step'],2), round(f['full_rectangle_ms_per_step'],2), f['rectangle_vs_half_identical'], f['stages_ms'], f['parity'])
print('pipelined', d['e2e'].get('pipelined'))
print('cpu', d['cpu_baseline'])
PY
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err
cut -c1-300 gpurun_out/r02z_bench_ref.json
tail -3 gpurun_out/r02z_bench_default.err
