#!/bin/bash
# round-2 GPU call 60: knobs of the host pipeline on one box (bytes in flight per upload SM, part schedules)
mkdir -p gpurun_out
S="WEALY_HOST_UP_THREADS=512;WEALY_HOST_UP_THREADS=256;WEALY_HOST_UP_THREADS=128;WEALY_HOST_UP_THREADS=256&WEALY_HOST_UP_SMS=12;WEALY_HOST_UP_THREADS=128&WEALY_HOST_UP_SMS=16"
S="$S;WEALY_HOST_CUM=0.1,0.25,0.5;WEALY_HOST_CUM=0.08,0.2,0.4;WEALY_HOST_CUM=0.15,0.4;WEALY_HOST_CUM=0.05,0.12,0.22,0.35,0.55;WEALY_HOST_CUM=0.1,0.2,0.35,0.55,0.8"
timeout 250 python tools/e2e_host_bench.py --parts 0 --steps 6 --sweep "$S" > gpurun_out/r02_host_knobs.json 2> gpurun_out/r02_host_knobs.err
tail -3 gpurun_out/r02_host_knobs.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_host_knobs.json').read().strip().splitlines()[-1])
for r in d['runs']: print(r['mode'], round(r['ms_per_step'],2))
for r in d['sweep']: print(r['rep'], round(r['ms_per_step'],2), r['identical'], r['setting'])
PY
