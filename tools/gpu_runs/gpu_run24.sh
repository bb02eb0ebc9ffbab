set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r24_all_tests.txt
