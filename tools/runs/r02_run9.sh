#!/bin/bash
# round-2 GPU call 9: staged sharded sweep (emulated), 4-chain drain
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02i_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02i_pytest.log | head -20
B="python bench.py --legs main --no-cpu --steps 20 --warmup 5"
run() { name=$1; shift; ( env "$@" timeout 300 $B ) > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02i_$name.json') if l.startswith('{')][-1])
    print('$name', 'value %.1f ms %.2f kernel %.2f e2e %.1f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['run']['map'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/r02i_$name.err').read()[-600:])
PY
}
run pair_a X=1
run pair_b X=1
run pair_t16 WEALY_TILES_PER_UNIT=16
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --sigma 4.0"
run hard_pair X=1
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --precision fp16"
run fp16_pair X=1
run fp16_single WEALY_SYM_PAIR=0
