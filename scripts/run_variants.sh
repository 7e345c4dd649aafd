#!/bin/bash
# on the GPU box: quick_bench every tuning build named on the command line (tags), fast mode, 65 536 envs
mkdir -p gpurun_out
for tag in "$@"; do
  lib=rl_rocket_b200/_C/librlrocket_$tag.so
  [ "$tag" = default ] && lib=rl_rocket_b200/_C/librlrocket.so
  if nm -D $lib | grep -q rlr_debug_phases; then RLR_LIB=$lib python scripts/phase_probe.py 2>&1 | tee -a gpurun_out/variants.log
  else RLR_LIB=$lib QB_SIZES=${QB_SIZES:-65536} QB_MODES=fast python scripts/quick_bench.py 2>&1 | tee -a gpurun_out/variants.log; fi
done
