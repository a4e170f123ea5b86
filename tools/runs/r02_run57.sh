#!/bin/bash
# round-2 GPU call 57: validation with wealy_eval_run_host as evaluate()'s route for pinned host embeddings
# (GPU suite, smoke, default bench), then the launch list of one pipelined evaluation
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02x_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^real" gpurun_out/r02x_pytest.log | head -20
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02x_smoke.log 2>&1
tail -5 gpurun_out/r02x_smoke.log | head -2
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02x_bench_default.json 2> gpurun_out/r02x_bench_default.err
tail -4 gpurun_out/r02x_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02x_bench_default.json') if l.startswith('{')][-1])
print('main value %.1f ms %.2f kernel %.2f frac %.3f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']))
e=d['e2e']; print('e2e %.1f (%.2f ms) route %s copy %s pipelined %.2f' % (e['value'], e['ms_per_step'], e['route'], e['copy_then_compute'] and round(e['copy_then_compute']['ms_per_step'],2), e['pipelined']['ms_per_step']))
print('parity', {k:v for k,v in d['parity'].items() if k!='note'}, e.get('abs_dMAP_vs_resident_path'))
PY
WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 0 --steps 2 --only-host 2>&1 >/dev/null | tail -10 > gpurun_out/r02x_host_trace.log
cat gpurun_out/r02x_host_trace.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_host_pipeline.csv python tools/e2e_host_bench.py --parts 0 --steps 1 --only-host > gpurun_out/r02x_ncu.log 2>&1
echo "ncu rc=$?"
