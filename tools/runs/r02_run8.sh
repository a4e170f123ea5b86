#!/bin/bash
# round-2 GPU call 8: fused loss call (no ATen launches), self-resetting unit counters; full suite + C4 + launch list of a loss step
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r02h_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02h_pytest.log | head -20
( time timeout 600 python bench.py --legs c4 --no-cpu ) > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench.json') if l.startswith('{')][-1])
print('main value %.1f ms %.2f kernel %.2f e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('c4', {k:(v['fwd_ms'],v['fwd_bwd_eager_ms'],v['fwd_bwd_graph_ms'],v['loss_rel_err_vs_cpu_fp32'],v.get('grad_rel_l2_vs_cpu_fp32')) for k,v in d['c4_loss'].items() if isinstance(v,dict)})
PY
cat > /tmp/loss_step.py <<PY
import sys, torch
sys.path.insert(0, '.')
import wealy_b200
from wealy_b200 import losses as wl
from wealy_b200.data import synth
s = synth.make_loss_batch(4096, 1024, seed=0, dtype=torch.bfloat16, device='cuda')
for mod in (wl.NTXentLoss(0.1), wl.CLEWSLoss()):
    z = s['z'].clone().requires_grad_(True)
    for _ in range(3):
        loss, _ = mod(s['label'], s['idx'], z); loss.backward()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push('step')
    loss, _ = mod(s['label'], s['idx'], z); loss.backward()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "step/" --csv --log-file gpurun_out/r02_launches_loss_step.csv python /tmp/loss_step.py > gpurun_out/r02h_ncu_loss.log 2>&1
echo "ncu rc=$?"; grep -c "gpu__time_duration" gpurun_out/r02_launches_loss_step.csv
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_loss_step.csv')) if len(r)>5]
i=next(k for k,r in enumerate(rows) if 'Kernel Name' in r); h=rows[i]; ki=h.index('Kernel Name'); mi=h.index('Metric Value')
for r in rows[i+1:]: print(r[mi], r[ki][:90])
PY
