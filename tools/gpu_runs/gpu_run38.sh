set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r38_all_tests.txt
