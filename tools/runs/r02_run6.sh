#!/bin/bash
# round-2 GPU call 6: pair kernel as default (full suite), C5 on the pair kernel, e2e after the upload-stream fix
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02f_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02f_pytest.log | head -20
( time timeout 900 python bench.py ) > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench_default.json') if l.startswith('{')][-1])
print('main value %.1f ms %.2f kernel %.2f frac %.3f e2e %.1f (%.2f ms) clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks']))
print('parity', d['parity'])
for k in ('c1_shs100k','c3_500k','c5_topk100'):
    c=d[k]; print(k, c['ms_per_step'], c.get('gpairs_per_s'), c.get('roofline_frac'), c.get('topk_path'), (c.get('parity') or c.get('parity_all_queries')))
print('c4', {k:(v['fwd_ms'],v['fwd_bwd_eager_ms'],v['fwd_bwd_graph_ms'],v['loss_rel_err_vs_cpu_fp32']) for k,v in d['c4_loss'].items() if isinstance(v,dict)})
PY
( WEALY_SYM_PAIR=0 timeout 600 python bench.py --legs c5 --no-cpu ) > gpurun_out/r02f_bench_c5_single.json 2> gpurun_out/r02f_bench_c5_single.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench_c5_single.json') if l.startswith('{')][-1]); c=d['c5_topk100']
print('c5 single-CTA', c['ms_per_step'], c['sweep_ms'])
PY
