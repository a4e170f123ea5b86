#!/bin/bash
# round-2 GPU call 10: where the pair kernel's MMA thread waits (diagnostic build), MMA floor, 2-chain drain re-check
mkdir -p gpurun_out
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3"
run() { name=$1; shift; ( env "$@" timeout 300 $B ) > gpurun_out/r02j_$name.json 2> gpurun_out/r02j_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02j_$name.json') if l.startswith('{')][-1])
    print('$name', 'value %.1f ms %.2f kernel %.2f e2e %.1f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['run']['map'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/r02j_$name.err').read()[-600:])
PY
}
run pair X=1
run pair2 X=1
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --sigma 0.05"
run floor_pair X=1
run floor_single WEALY_SYM_PAIR=0
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --sigma 4.0"
run hard_pair X=1
B="python bench.py --legs main --no-cpu --steps 3 --warmup 3"
( WEALY_LIB=$PWD/audio-based-lyrics-matching_b200/lib/libwealy_b200_prof.so timeout 300 $B ) > gpurun_out/r02j_prof.log 2>&1
grep "wealy waits" gpurun_out/r02j_prof.log | tail -8
( WEALY_LIB=$PWD/audio-based-lyrics-matching_b200/lib/libwealy_b200_prof.so timeout 300 $B --sigma 0.05 ) > gpurun_out/r02j_prof_floor.log 2>&1
grep "wealy waits" gpurun_out/r02j_prof_floor.log | tail -4
