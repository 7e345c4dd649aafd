#!/bin/bash
# One lighter GPU-box visit: parity tests, smoke, contract bench (K = 500, K = 20, reference arm), ncu launch list, one --set full capture of the chained step kernel.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_steps20.json 2>> $O/${TAG}_bench.err; echo "bench20 exit $?"
python bench.py --impl reference --steps 3 --warmup 3 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2300 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:step_kernel<rlr_fast::Sys6, .bool.1>" -s 1965 -c 1 -f \
    -o $O/${TAG}_stepchained python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_ncu_stepchained.log 2>&1; echo "ncu chained exit $?"
ls -la $O | grep ${TAG}
