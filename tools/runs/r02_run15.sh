#!/bin/bash
# round-2 GPU call 15: ncu captures of the C5 kernels, exported to CSV on the box (the reports themselves exceed the merge limit);
# scheduling-group sweep for the pair kernel
mkdir -p gpurun_out /tmp/rep
CMD5="python tools/c5_once.py"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 2 -c 1 -f -o /tmp/rep/c5 $CMD5 > gpurun_out/r02n_ncu_c5.log 2>&1
echo "c5 rc=$?"
ncu -i /tmp/rep/c5.ncu-rep --page raw --csv > gpurun_out/r02_ncu_topk_c5_raw.csv 2>/dev/null
ncu -i /tmp/rep/c5.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/c5_src.csv 2>/dev/null
python tools/src_lines.py /tmp/rep/c5_src.csv > gpurun_out/r02_src_lines_topk_c5.txt 2>&1
ncu --set full --clock-control none -k regex:GroupMaxEpi -s 2 -c 1 -f -o /tmp/rep/c5b $CMD5 > gpurun_out/r02n_ncu_c5b.log 2>&1
ncu -i /tmp/rep/c5b.ncu-rep --page raw --csv > gpurun_out/r02_ncu_topk_prepass_raw.csv 2>/dev/null
B="python bench.py --legs main --no-cpu --steps 10 --warmup 3"
run() { name=$1; shift; ( env "$@" timeout 300 $B ) > gpurun_out/r02n_$name.json 2> gpurun_out/r02n_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02n_$name.json') if l.startswith('{')][-1])
    print('$name', 'value %.1f ms %.2f kernel %.2f e2e %.1f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/r02n_$name.err').read()[-600:])
PY
}
run base X=1
run gr20 WEALY_GROUP_ROWS=20
run gr74 WEALY_GROUP_ROWS=74
run gr148 WEALY_GROUP_ROWS=148
run rounds32 WEALY_ROUNDS=32
run t4 WEALY_TILES_PER_UNIT=4
