#!/bin/bash
# round-2 GPU call 14: ncu evidence for profiles/ (launch lists + full captures of the final kernels)
mkdir -p gpurun_out
CMD="python bench.py --legs main --no-cpu --steps 2 --warmup 3"
$CMD > gpurun_out/r02m_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_fp16x3.csv $CMD > gpurun_out/r02m_ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_sweep_pair $CMD > gpurun_out/r02m_ncu_full.log 2>&1
echo "full rc=$?"
CMD5="python tools/c5_once.py"
cat > tools/c5_once.py <<PY
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wealy_b200
from wealy_b200 import evaluation as we
from wealy_b200.data import synth
s = synth.make_eval_set(50_000, 2048, seed=5, dist="lyric_covers_test", device="cuda", md5_ids=False)
plan = we.EvalPlan(s["c"], s["i"], s["c"], s["i"])
for _ in range(4):
    r = plan.run(s["z"], s["z"], topk=100)
torch.cuda.synchronize()
print("path", plan.last_topk_path(), plan.stage_ms())
PY
$CMD5 > gpurun_out/r02m_c5_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_topk_c5.csv $CMD5 > gpurun_out/r02m_ncu_list5.log 2>&1
echo "list5 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_topk_c5 $CMD5 > gpurun_out/r02m_ncu_full5.log 2>&1
echo "full5 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:GroupMaxEpi -s 2 -c 1 -f -o gpurun_out/r02_prof_topk_prepass $CMD5 > gpurun_out/r02m_ncu_full5b.log 2>&1
echo "full5b rc=$?"
tail -2 gpurun_out/r02m_c5_plain.log
