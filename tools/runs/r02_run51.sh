#!/bin/bash
# round-2 GPU call 51: ncu launch list + full capture of the headline kernel with the final scheduling defaults
mkdir -p gpurun_out /tmp/rep
CMD="python bench.py --legs main --no-cpu --steps 2 --warmup 3"
$CMD > gpurun_out/r02fin_plain.log 2>&1 || { tail -5 gpurun_out/r02fin_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_fp16x3_final.csv $CMD > gpurun_out/r02fin_ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 4 -c 1 -f -o /tmp/rep/main $CMD > gpurun_out/r02fin_ncu_main.log 2>&1
echo "main rc=$?"
ncu -i /tmp/rep/main.ncu-rep --page raw --csv > gpurun_out/r02_ncu_sweep_pair_final_raw.csv 2>/dev/null
ncu -i /tmp/rep/main.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/main_src.csv 2>/dev/null
python tools/src_lines.py /tmp/rep/main_src.csv > gpurun_out/r02_src_lines_sweep_pair_final.txt 2>&1
python tools/ncu_raw_metrics.py gpurun_out/r02_ncu_sweep_pair_final_raw.csv
