N=$1
python tools/dp_loss_bench.py 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/dp_loss_bench.py 2>&1 | tail -1
