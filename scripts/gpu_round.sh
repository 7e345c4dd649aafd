#!/bin/bash
# One GPU-box visit: parity tests, contract bench (both arms), ncu launch list and one --set full capture of every kernel of DESIGN.md section 3.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt
if [ -z "$ONLY_NCU" ]; then
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
tail -c 1500 $O/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 3 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref exit $?"
cat $O/${TAG}_bench_ref.json
fi
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
if [ -z "$ONLY_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list exit $?"
fi
cap() {   # name, kernel regex, skip
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $O/${TAG}_$1 $B > $O/${TAG}_ncu_$1.log 2>&1; echo "ncu $1 exit $?"
}
cap step "step_kernel<rlr_fast::Sys6>" 1965
cap rollout "rollout_kernel<rlr_fast::Sys6>" 1
cap step3 "step_kernel<rlr_fast::Sys3>" 160
cap step3dense "step_kernel<rlr_fast::Sys3Dense>" 160
cap reset "reset_kernel<rlr_exact::Sys6>" 0
cap stepf32 "step_kernel_f32" 160
ls -la $O | tail -20
