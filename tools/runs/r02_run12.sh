#!/bin/bash
# round-2 GPU call 12 (8 GPUs): bench at N=8 (weak-scaling headline + 500k strong-scaling leg + parity), data-parallel losses at 8 ranks
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02k_smi.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r02k_bench_8gpu.json 2> gpurun_out/r02k_bench_8gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02k_bench_8gpu.json') if l.startswith('{')][-1])
    print('N=8 value %.1f ms %.2f kernel %.2f e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step']))
    print('parity', {k:v for k,v in d.get('parity',{}).items() if k!='note'})
    c=d['c3_500k']; print('c3', c['ms_per_step'], c['gpairs_per_s'], c['roofline_frac'], c['parity']['item_ranks_out_of_band'], c['parity']['abs_dMAP'])
except Exception as e:
    print('bench8 FAILED', e); print(open('gpurun_out/r02k_bench_8gpu.err').read()[-2500:])
PY
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tools/dp_loss_bench.py ) > gpurun_out/r02k_dp_loss_8gpu.json 2> gpurun_out/r02k_dp_loss_8gpu.err
tail -1 gpurun_out/r02k_dp_loss_8gpu.json | cut -c1-1500
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 10 --warmup 3 --legs main ) > gpurun_out/r02k_bench_4gpu.json 2> gpurun_out/r02k_bench_4gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02k_bench_4gpu.json') if l.startswith('{')][-1])
    print('N=4 value %.1f ms %.2f kernel %.2f e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step']))
except Exception as e:
    print('bench4 FAILED', e); print(open('gpurun_out/r02k_bench_4gpu.err').read()[-1500:])
PY
