set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_full_size_gpu.py -q -m gpu -s 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r28_full_size.txt
