#!/bin/bash
# round-2 GPU call 33: A/B of the early accumulator release in the pair kernel's dynamic-chunk epilogue
# (the "late" library: nvcc ... -DWEALY_LATE_RELEASE -o audio-based-lyrics-matching_b200/lib/libwealy_b200_late.so audio-based-lyrics-matching_b200/csrc/api.cu, same flags as build.py; not kept)
mkdir -p gpurun_out
L=$PWD/audio-based-lyrics-matching_b200/lib
run() {  # label, extra env, extra args
  ( env $2 timeout 600 python bench.py --legs main,c5 --no-cpu --steps 10 --warmup 3 $3 ) > gpurun_out/r02ab_$1.json 2> gpurun_out/r02ab_$1.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02ab_$1.json') if l.startswith('{')][-1])
print('$1 value %.1f ms %.2f kernel %.2f map %.6f clk %s | c5 %.2f ms sweep %.2f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['run']['map'], d['clocks']['sm_mhz'], d['c5_topk100']['ms_per_step'], d['c5_topk100']['sweep_ms']))
PY
}
run early ""
run late "WEALY_LIB=$L/libwealy_b200_late.so"
run early2 ""
run late2 "WEALY_LIB=$L/libwealy_b200_late.so"
run hard_early "X=1" "--sigma 4.0"
run hard_late "WEALY_LIB=$L/libwealy_b200_late.so" "--sigma 4.0"
( timeout 900 python -m pytest tests/test_gpu_eval.py -x -q ) 2>&1 | tail -2
