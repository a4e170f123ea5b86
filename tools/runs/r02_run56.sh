#!/bin/bash
# round-2 GPU call 56: how many SMs the upload kernel needs (WEALY_HOST_UP_SMS), part schedules on top, traces
mkdir -p gpurun_out
for u in 2 4 6 8 12; do
  echo "== upload on $u SMs"
  WEALY_HOST_UP_SMS=$u timeout 200 python tools/e2e_host_bench.py --parts 0,7 --steps 4 2>gpurun_out/r02k_err_$u.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['runs']: print(r['mode'], round(r['ms_per_step'],2), r['identical_to_copy_path'])
"
  tail -2 gpurun_out/r02k_err_$u.log
done
for u in 4 8; do
echo "== trace, $u SMs, 5 parts"
WEALY_HOST_UP_SMS=$u WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 0 --steps 2 2>&1 >/dev/null | tail -5
done
