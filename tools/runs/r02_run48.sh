#!/bin/bash
# round-2 GPU call 48: scheduling knobs of the symmetric sweep on the final kernel (tiles per unit, group rows)
mkdir -p gpurun_out
run() {
  ( env $2 timeout 600 python bench.py --legs main --no-cpu --steps 10 --warmup 3 ) > gpurun_out/r02tune_$1.json 2> gpurun_out/r02tune_$1.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02tune_$1.json') if l.startswith('{')][-1])
print('$1 value %.1f ms %.2f kernel %.2f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))
PY
}
run gr56 "WEALY_GROUP_ROWS=56"
run gr74 "WEALY_GROUP_ROWS=74"
run gr100 "WEALY_GROUP_ROWS=100"
run gr120 "WEALY_GROUP_ROWS=120"
run gr74_tpu16 "WEALY_GROUP_ROWS=74 WEALY_TILES_PER_UNIT=16"
run base3 "X=1"
run gr74b "WEALY_GROUP_ROWS=74"
run gr100b "WEALY_GROUP_ROWS=100"
