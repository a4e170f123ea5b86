#!/bin/bash
# round-2 GPU call 44: the documented environment knobs still select working kernels (evaluation tests under each)
mkdir -p gpurun_out
for cfg in "WEALY_PAIR_INTERLEAVE=0" "WEALY_SYM_PAIR=0" "WEALY_RECT_PAIR=0" "WEALY_PAIR_DYN=0" "WEALY_PAIR_EPI_WARPS=8" "WEALY_SYM_LEVELS=3" "WEALY_SYM_LEVELS=2" "WEALY_SYM_TRACKS=0" "WEALY_SYM_TOPK=0" "WEALY_POOL_KEEP_MB=0"; do
  ( env $cfg timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_eval_chunked.py -x -q -k "not equals_rectangle" ) > gpurun_out/r02knob.log 2>&1
  echo "$cfg: $(tail -1 gpurun_out/r02knob.log)"
  grep -E "^(FAILED|ERROR)" gpurun_out/r02knob.log | head -3
done
