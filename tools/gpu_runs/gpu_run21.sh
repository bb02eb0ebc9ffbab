set -x
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ibm_slabs_gpu.py -q -m gpu -k "peer_mapped_slabs_with and 2-1-7" 2>&1 | grep -E "Error|passed|failed" | tail -5 ) 2>&1 | tee gpurun_out/r21_a.txt
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r21_all_tests.txt
