set -x
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ibm_slabs_gpu.py -q -m gpu -k "peer_mapped_slabs_with and 2-1-7" 2>&1 | grep -E "Error|passed|failed" | tail -5 ) 2>&1 | tee gpurun_out/r20_a.txt
( LBM_B200_OVERLAP=0 timeout 300 python -m pytest tests/test_ibm_slabs_gpu.py -q -m gpu -k "peer_mapped_slabs_with and 2-1-7" 2>&1 | grep -E "Error|passed|failed" | tail -5 ) 2>&1 | tee gpurun_out/r20_b.txt
( CUDA_MODULE_LOADING=EAGER timeout 300 python -m pytest tests/test_ibm_slabs_gpu.py -q -m gpu -k "peer_mapped_slabs_with and 2-1-7" 2>&1 | grep -E "Error|passed|failed" | tail -5 ) 2>&1 | tee gpurun_out/r20_c.txt
( CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python -m pytest tests/test_ibm_slabs_gpu.py -q -m gpu -k "peer_mapped_slabs_with and 2-1-7" 2>&1 | grep -E "Error|passed|failed" | tail -5 ) 2>&1 | tee gpurun_out/r20_d.txt
