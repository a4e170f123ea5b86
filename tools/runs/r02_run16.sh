#!/bin/bash
# round-2 GPU call 16: ncu captures of the C5 kernels, exported to CSV on the box (the reports exceed the merge limit)
mkdir -p gpurun_out /tmp/rep
CMD5="python tools/c5_once.py"
$CMD5 > gpurun_out/r02o_c5_plain.log 2>&1 || { tail -5 gpurun_out/r02o_c5_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 2 -c 1 -f -o /tmp/rep/c5 $CMD5 > gpurun_out/r02o_ncu_c5.log 2>&1
echo "c5 rc=$?"
ncu -i /tmp/rep/c5.ncu-rep --page raw --csv > gpurun_out/r02_ncu_topk_c5_raw.csv 2>/dev/null
ncu -i /tmp/rep/c5.ncu-rep --page source --csv --print-source cuda,sass > /tmp/rep/c5_src.csv 2>/dev/null
python tools/src_lines.py /tmp/rep/c5_src.csv > gpurun_out/r02_src_lines_topk_c5.txt 2>&1
ncu --set full --clock-control none -k regex:GroupMaxEpi -s 2 -c 1 -f -o /tmp/rep/c5b $CMD5 > gpurun_out/r02o_ncu_c5b.log 2>&1
echo "c5b rc=$?"
ncu -i /tmp/rep/c5b.ncu-rep --page raw --csv > gpurun_out/r02_ncu_topk_prepass_raw.csv 2>/dev/null
ls -la gpurun_out/r02_ncu_topk*; head -4 gpurun_out/r02_src_lines_topk_c5.txt
