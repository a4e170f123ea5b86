#!/bin/bash
# round-2 GPU call 35: full GPU suite on the 4-level default, intermediate hardness A/B, ncu captures of the new default
# headline kernel and of the chunked half sweep (exported to CSV on the box)
mkdir -p gpurun_out /tmp/rep
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02q_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02q_pytest.log | head
run() {
  ( env $2 timeout 600 python bench.py --legs main --no-cpu --steps 10 --warmup 3 $3 ) > gpurun_out/r02q_$1.json 2> gpurun_out/r02q_$1.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02q_$1.json') if l.startswith('{')][-1])
print('$1 value %.1f ms %.2f kernel %.2f map %.6f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['run']['map'], d['clocks']['sm_mhz']))
PY
}
run mid_lv3 "WEALY_SYM_LEVELS=3" "--sigma 2.5"
run mid_lv4 "WEALY_SYM_LEVELS=4" "--sigma 2.5"
run mid3_lv3 "WEALY_SYM_LEVELS=3" "--sigma 3.2"
run mid3_lv4 "WEALY_SYM_LEVELS=4" "--sigma 3.2"
CMD="python bench.py --legs main --no-cpu --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_fp16x3_lv4.csv $CMD > gpurun_out/r02q_ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 4 -c 1 -f -o /tmp/rep/main $CMD > gpurun_out/r02q_ncu_main.log 2>&1
echo "main rc=$?"
ncu -i /tmp/rep/main.ncu-rep --page raw --csv > gpurun_out/r02_ncu_sweep_pair_lv4_raw.csv 2>/dev/null
ncu -i /tmp/rep/main.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/main_src.csv 2>/dev/null
python tools/src_lines.py /tmp/rep/main_src.csv > gpurun_out/r02_src_lines_sweep_pair_lv4.txt 2>&1
CMDF="python tools/f1_once.py"
$CMDF > gpurun_out/r02q_f1_plain.log 2>&1; tail -1 gpurun_out/r02q_f1_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_chunked_f1.csv $CMDF > gpurun_out/r02q_ncu_f1list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 2 -c 1 -f -o /tmp/rep/f1 $CMDF > gpurun_out/r02q_ncu_f1.log 2>&1
echo "f1 rc=$?"
ncu -i /tmp/rep/f1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_chunked_f1_raw.csv 2>/dev/null
ncu -i /tmp/rep/f1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/f1_src.csv 2>/dev/null
python tools/src_lines.py /tmp/rep/f1_src.csv > gpurun_out/r02_src_lines_chunked_f1.txt 2>&1
ls -la gpurun_out/r02_ncu_sweep_pair_lv4_raw.csv gpurun_out/r02_ncu_chunked_f1_raw.csv; head -8 gpurun_out/r02_src_lines_sweep_pair_lv4.txt; head -8 gpurun_out/r02_src_lines_chunked_f1.txt
