for sg in 2.4 1.0; do for lv in 3 4; do WEALY_SYM_LEVELS=$lv timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 $sg 2>&1 | tail -1; done; done
echo backoff
for sg in 2.4 1.0; do WEALY_LIB=$PWD/audio-based-lyrics-matching_b200/lib/libwealy_b200_backoff.so WEALY_SYM_LEVELS=3 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 $sg 2>&1 | tail -1; done
WEALY_SYM=0 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 1.0 2>&1 | tail -1
