#!/bin/bash
# round-2 GPU call 7 (2 GPUs): NCCL tests, bench at N=2 (headline + c3 strong-scaling leg + parity), data-parallel loss overlap
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02g_smi.txt
( time timeout 900 python -m pytest tests/test_gpu_dist_eval.py tests/test_gpu_dist_losses.py -q ) > gpurun_out/r02g_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|skipped" gpurun_out/r02g_pytest.log | head
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02g_bench_2gpu.json') if l.startswith('{')][-1])
    print('N=2 value %.1f ms %.2f kernel %.2f e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step']))
    print('parity', d.get('parity'))
    c=d['c3_500k']; print('c3', c['ms_per_step'], c['gpairs_per_s'], c['roofline_frac'], c['parity']['item_ranks_out_of_band'], c['parity']['abs_dMAP'])
except Exception as e:
    print('bench2 FAILED', e); print(open('gpurun_out/r02g_bench_2gpu.err').read()[-2500:])
PY
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/dp_loss_bench.py ) > gpurun_out/r02g_dp_loss_2gpu.json 2> gpurun_out/r02g_dp_loss_2gpu.err
tail -1 gpurun_out/r02g_dp_loss_2gpu.json | cut -c1-1500
tail -3 gpurun_out/r02g_dp_loss_2gpu.err
