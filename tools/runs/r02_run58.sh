#!/bin/bash
# round-2 GPU call 58: validation with the early upload inside the plan build (GPU suite, smoke, default bench),
# early upload on / off, launch list of the pipelined evaluation
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r02w_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^real" gpurun_out/r02w_pytest.log | head -20
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02w_smoke.log 2>&1
tail -5 gpurun_out/r02w_smoke.log | head -2
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02w_bench_default.json 2> gpurun_out/r02w_bench_default.err
tail -3 gpurun_out/r02w_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02w_bench_default.json') if l.startswith('{')][-1])
print('main value %.1f ms %.2f kernel %.2f frac %.3f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']))
e=d['e2e']; print('e2e %.1f (%.2f ms) route %s copy %s pipelined %.2f' % (e['value'], e['ms_per_step'], e['route'], e['copy_then_compute'] and round(e['copy_then_compute']['ms_per_step'],2), e['pipelined']['ms_per_step']))
print('parity', {k:v for k,v in d['parity'].items() if k!='note'}, e.get('abs_dMAP_vs_resident_path'))
PY
for early in 1 0 1 0; do
  echo "== early upload $early"
  WEALY_HOST_EARLY=$early timeout 200 python tools/e2e_host_bench.py --parts 0 --steps 6 --only-host 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['runs']: print(r['mode'], round(r['ms_per_step'],2), r['map'])
"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prep_rows_stream|gemm_pair|pos_pairs|pos_sort|ap_reduce|segment_lookup" -c 80 --csv --log-file gpurun_out/r02_launches_host_pipeline.csv python tools/e2e_host_bench.py --parts 0 --steps 1 --only-host > gpurun_out/r02w_ncu.log 2>&1
echo "ncu rc=$?"
