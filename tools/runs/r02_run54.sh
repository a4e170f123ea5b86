#!/bin/bash
# round-2 GPU call 54: where the upload / sweep pipeline of wealy_eval_run_host loses time (per-part time stamps), upload kernel geometry
mkdir -p gpurun_out
for cfg in "64 1" "32 1" "32 2" "64 2"; do
  set -- $cfg
  echo "== upload threads $1 grid x$2"
  WEALY_HOST_UP_THREADS=$1 WEALY_HOST_UP_GRID=$2 timeout 200 python tools/e2e_host_bench.py --parts 0,2 --steps 4 2>gpurun_out/r02i_err_$1_$2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['runs']: print(r['mode'], round(r['ms_per_step'],2), r['identical_to_copy_path'])
"
done
echo "== trace, 5 parts"
WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 0 --steps 2 2>&1 >/dev/null | tail -12
echo "== trace, 2 parts"
WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 2 --steps 2 2>&1 >/dev/null | tail -6
echo "== trace, 1 part"
WEALY_HOST_TRACE=1 timeout 200 python tools/e2e_host_bench.py --parts 1 --steps 2 2>&1 >/dev/null | tail -3
