#!/bin/bash
# One GPU-box visit: parity tests, contract bench, ncu launch list and one --set full capture of the two hot kernels.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $O/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 3 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref exit $?"
cat $O/${TAG}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 165 -c 1 -f -o $O/${TAG}_step \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_ncu_step.log 2>&1; echo "ncu step exit $?"
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o $O/${TAG}_rollout \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_ncu_rollout.log 2>&1; echo "ncu rollout exit $?"
ls -la $O
