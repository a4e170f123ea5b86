#!/bin/bash
# round-2 GPU call 53: wealy_eval_run_host (upload / prep / sweep pipeline): parity tests, part schedules vs copy-then-compute
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_eval_host.py -x -q --durations=5 ) > gpurun_out/r02h_pytest.log 2>&1
tail -15 gpurun_out/r02h_pytest.log
timeout 300 python tools/e2e_host_bench.py > gpurun_out/r02h_e2e.json 2> gpurun_out/r02h_e2e.err
tail -3 gpurun_out/r02h_e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02h_e2e.json').read().strip().splitlines()[-1])
print("plan_build_ms", d.get("plan_build_ms"))
for r in d["runs"]: print(r)
PY
