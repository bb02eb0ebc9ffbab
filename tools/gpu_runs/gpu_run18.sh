set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 600 python -m pytest tests/test_ibm_slabs_gpu.py tests/test_restart_validation_gpu.py -q -m gpu 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r18_new_tests.txt
( time timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_ibm_slabs_gpu.py --deselect tests/test_restart_validation_gpu.py 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r18_all_tests.txt
LBM_B200_OVERLAP=0 python tools/config_bench.py c5 c5d c3 c3l c2 --steps 200 2>&1 | grep '^{' | tee gpurun_out/r18_config_bench_serial.txt
LBM_B200_OVERLAP=1 python tools/config_bench.py c5 c5d c3 c3l c2 --steps 200 2>&1 | grep '^{' | tee gpurun_out/r18_config_bench_overlap.txt
