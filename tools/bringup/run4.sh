L=$PWD/audio-based-lyrics-matching_b200/lib
for v in "" _hint _lane0 _both ""; do
  for sg in 2.4 1.0; do echo "variant '$v'"; WEALY_LIB=$L/libwealy_b200$v.so WEALY_SYM_LEVELS=3 timeout 120 python tools/gpu_diag.py time fp16x3 100000 1024 0 $sg 2>&1 | tail -1; done
done
