set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "run_from_host" 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r35_tests.txt
