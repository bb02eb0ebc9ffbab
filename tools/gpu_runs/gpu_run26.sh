set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_shim_gpu.py tests/test_parity_gpu.py -q -m gpu 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r26_tests.txt
cd examples/_bin
for b in ex_c1_tg_256 ex_c2_pois_1024x256; do ./$b --steps 4000 --save-int 4000 --fast | grep SHIM; LBM_B200_GRAPH=0 ./$b --steps 4000 --save-int 4000 --fast | grep SHIM; ./$b --steps 4000 --save-int 4000 | grep SHIM; done 2>&1 | tee ../../gpurun_out/r26_shim_small.txt
