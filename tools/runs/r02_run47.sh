#!/bin/bash
# round-2 GPU call 47: scheduling knobs of the symmetric sweep on the final kernel (tiles per unit, group rows)
mkdir -p gpurun_out
run() {
  ( env $2 timeout 600 python bench.py --legs main --no-cpu --steps 10 --warmup 3 ) > gpurun_out/r02tune_$1.json 2> gpurun_out/r02tune_$1.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02tune_$1.json') if l.startswith('{')][-1])
print('$1 value %.1f ms %.2f kernel %.2f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))
PY
}
run base "X=1"
run tpu4 "WEALY_TILES_PER_UNIT=4"
run tpu16 "WEALY_TILES_PER_UNIT=16"
run tpu32 "WEALY_TILES_PER_UNIT=32"
run gr9 "WEALY_GROUP_ROWS=18"
run gr37 "WEALY_GROUP_ROWS=74"
run gr148 "WEALY_GROUP_ROWS=296"
run base2 "X=1"
