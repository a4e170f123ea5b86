for il in 1 0; do
echo interleave=$il
WEALY_PAIR_INTERLEAVE=$il WEALY_SYM_PAIR=1 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 1.0 2>&1 | tail -1
WEALY_PAIR_INTERLEAVE=$il WEALY_SYM_PAIR=1 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
WEALY_PAIR_INTERLEAVE=$il WEALY_SYM_PAIR=1 timeout 120 python tools/gpu_diag.py time fp16 100000 1024 2>&1 | tail -1
done
timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 1.0 2>&1 | tail -1
