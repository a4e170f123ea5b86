#!/bin/bash
# round-2 GPU call 17: 16 epilogue warps on the pair kernel (correctness + timing)
mkdir -p gpurun_out
( WEALY_PAIR_EPI_WARPS=16 timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -k "symmetric_sweep or parity_with_oracle or config1 or sharded or kat or ragged" ) > gpurun_out/r02p_pytest16.log 2>&1
tail -2 gpurun_out/r02p_pytest16.log
B="python bench.py --legs main --no-cpu --steps 20 --warmup 5"
run() { name=$1; shift; ( env "$@" timeout 300 $B ) > gpurun_out/r02p_$name.json 2> gpurun_out/r02p_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02p_$name.json') if l.startswith('{')][-1])
    print('$name', 'value %.1f ms %.2f kernel %.2f e2e %.1f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['run']['map'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/r02p_$name.err').read()[-600:])
PY
}
run w12 X=1
run w16 WEALY_PAIR_EPI_WARPS=16
run w12b X=1
run w16b WEALY_PAIR_EPI_WARPS=16
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --sigma 4.0"
run hard_w12 X=1
run hard_w16 WEALY_PAIR_EPI_WARPS=16
