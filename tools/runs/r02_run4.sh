#!/bin/bash
# round-2 GPU call 4: pair kernel variants (dynamic chunk claiming), correctness under the pair kernel, ncu captures
mkdir -p gpurun_out
B="python bench.py --legs main --no-cpu --steps 20 --warmup 5"
run() { name=$1; shift; ( env "$@" timeout 300 $B ) > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02d_$name.json') if l.startswith('{')][-1])
    print('$name', 'value %.1f ms %.2f kernel %.2f e2e %.1f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['run']['map'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/r02d_$name.err').read()[-600:])
PY
}
( WEALY_SYM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -k "symmetric_sweep or parity_with_oracle or config1 or sharded or kat or ragged" ) > gpurun_out/r02d_pytest_pair.log 2>&1
tail -3 gpurun_out/r02d_pytest_pair.log
run single WEALY_SYM_PAIR=0
run pair_static12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=0
run pair_dyn12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1
run pair_dyn8 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1 WEALY_PAIR_EPI_WARPS=8
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3 --sigma 4.0"
run hard_single WEALY_SYM_PAIR=0
run hard_pair_dyn12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=1
run hard_pair_static12 WEALY_SYM_PAIR=1 WEALY_PAIR_DYN=0
# ncu: launch list of the default bench command, then one full capture of the pair kernel with source correlation
CMD="python bench.py --legs main --no-cpu --steps 2 --warmup 3"
WEALY_SYM_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 4 -c 1 -f -o gpurun_out/r02_prof_pair $CMD > gpurun_out/r02d_ncu_pair.log 2>&1
echo "ncu pair rc=$?"
