for i in $(seq 1 12); do timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "chained" 2>&1 | tail -1; done
