#!/bin/bash
# round-2 GPU call 55: upload kernel on SMs of its own (WEALY_HOST_UP_SMS) instead of next to the sweep's CTAs
mkdir -p gpurun_out
for u in 8 16 24; do
  echo "== upload on $u SMs"
  WEALY_HOST_UP_SMS=$u timeout 200 python tools/e2e_host_bench.py --parts 0,7 --steps 4 2>gpurun_out/r02j_err_$u.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['runs']: print(r['mode'], round(r['ms_per_step'],2), r['identical_to_copy_path'])
"
  tail -2 gpurun_out/r02j_err_$u.log
done
echo "== trace, 16 SMs, 5 parts"
WEALY_HOST_UP_SMS=16 WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 0 --steps 2 2>&1 >/dev/null | tail -5
echo "== trace, 8 SMs, 1 part"
WEALY_HOST_UP_SMS=8 WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 1 --steps 2 2>&1 >/dev/null | tail -1
