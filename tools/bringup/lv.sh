for lv in 2 3 4; do
WEALY_SYM_LEVELS=$lv python tools/gpu_diag.py time fp16 100000 1024 2>&1 | tail -1
WEALY_SYM_LEVELS=$lv python tools/gpu_diag.py time fp16x3 100000 1024 2>&1 | tail -1
done
