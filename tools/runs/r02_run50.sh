#!/bin/bash
# round-2 GPU call 50: scheduling group size for the top-k sweep (C5) and the chunked half sweep (f1)
mkdir -p gpurun_out
run() {
  ( env $2 timeout 600 python bench.py --legs main,c5,f1 --no-cpu --steps 3 --warmup 3 ) > gpurun_out/r02grp2_$1.json 2> gpurun_out/r02grp2_$1.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02grp2_$1.json') if l.startswith('{')][-1])
print('$1 c5 %.2f ms sweep %.2f | f1 %.2f ms sweep %.2f' % (d['c5_topk100']['ms_per_step'], d['c5_topk100']['sweep_ms'], d['f1_chunked']['ms_per_step'], d['f1_chunked']['sweep_ms']))
PY
}
run default "X=1"
run gr74 "WEALY_GROUP_ROWS=74"
run default2 "X=1"
run gr74b "WEALY_GROUP_ROWS=74"
run gr56 "WEALY_GROUP_ROWS=56"
