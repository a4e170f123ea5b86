N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
tail -1 gpurun_out/bench_n$N.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --tracks 500000 > gpurun_out/bench_c3_n$N.log 2>&1
tail -1 gpurun_out/bench_c3_n$N.log | cut -c1-200
