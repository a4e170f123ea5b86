#!/bin/bash
# round-2 GPU call 34: dense threshold levels of the pair kernel's epilogue (2 / 3 / 4)
mkdir -p gpurun_out
run() {
  ( env $2 timeout 600 python bench.py --legs main --no-cpu --steps 10 --warmup 3 $3 ) > gpurun_out/r02lv_$1.json 2> gpurun_out/r02lv_$1.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02lv_$1.json') if l.startswith('{')][-1])
print('$1 value %.1f ms %.2f kernel %.2f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['run']['map'], d['clocks']['sm_mhz']))
PY
}
run lv3 "WEALY_SYM_LEVELS=3"
run lv2 "WEALY_SYM_LEVELS=2"
run lv4 "WEALY_SYM_LEVELS=4"
run lv3b "WEALY_SYM_LEVELS=3"
run lv2b "WEALY_SYM_LEVELS=2"
run hard_lv3 "WEALY_SYM_LEVELS=3" "--sigma 4.0"
run hard_lv2 "WEALY_SYM_LEVELS=2" "--sigma 4.0"
run hard_lv4 "WEALY_SYM_LEVELS=4" "--sigma 4.0"
