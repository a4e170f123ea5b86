#!/bin/bash
# round-2 GPU call 21: EvalPipeline, smoke, memcheck attempt, bench e2e fields
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_eval.py -x -q -k "pipeline or kat or topk_matches" ) > gpurun_out/r02s_pytest.log 2>&1
tail -3 gpurun_out/r02s_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02s_smoke.log 2>&1; tail -1 gpurun_out/r02s_smoke.log
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02s_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02s_memcheck.log
( timeout 600 python bench.py --legs main --no-cpu --steps 10 --warmup 3 ) > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02s_bench.json') if l.startswith('{')][-1])
print('value %.1f  e2e %.1f (%.2f ms)  pipelined %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], {k:v for k,v in d['e2e']['pipelined'].items() if k!='api'}))
PY
tail -3 gpurun_out/r02s_bench.err
