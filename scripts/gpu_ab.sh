#!/bin/bash
# One GPU-box visit for an A/B of tuning builds: GPU parity tests on the default build, then quick_bench / phase probe / warp trace.
# usage (under gpurun): bash scripts/gpu_ab.sh "<tags for quick_bench>" [pytest-args]
O=gpurun_out; mkdir -p $O
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q ${2:-} > $O/ab_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/ab_pytest.log
fi
rm -f $O/variants.log
for tag in $1; do
  lib=rl_rocket_b200/_C/librlrocket_$tag.so
  [ "$tag" = default ] && lib=rl_rocket_b200/_C/librlrocket.so
  RLR_LIB=$lib QB_SIZES=${QB_SIZES:-65536} QB_MODES=fast python scripts/quick_bench.py 2>&1 | grep "^\[" | tee -a $O/variants.log
done
[ -f rl_rocket_b200/_C/librlrocket_PT.so ] && RLR_LIB=rl_rocket_b200/_C/librlrocket_PT.so python scripts/phase_probe.py 2>&1 | tee $O/phase_probe.txt
[ -f rl_rocket_b200/_C/librlrocket_WT.so ] && RLR_LIB=rl_rocket_b200/_C/librlrocket_WT.so python scripts/warp_trace.py 2>&1 | tee $O/warp_trace.txt
