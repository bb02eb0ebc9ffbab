# header-shim drivers on the GPU: parity against the reference's CUDA solver + the BASELINE configs through the C++ path
set -x
mkdir -p gpurun_out
python -m pytest tests/test_shim_gpu.py -q -s 2>&1 | tail -40
cd /tmp
for b in "ex_c3_lid_4096 --steps 200 --save-int 200 --fast" "ex_c5_cyl_8192x2048 --steps 200 --save-int 200 --fast" "ex_c2_pois_1024x256 --steps 2000 --save-int 2000 --fast" "ex_c4_tg_32768 --steps 40 --save-int 40 --fast" "ex_c4_tg_32768 --steps 40 --save-int 40"; do
  set -- $b; /root/repo/examples/_bin/$@ | grep -E "SHIM_RESULT|error" ; done 2>&1 | tee /root/repo/gpurun_out/shim_configs.txt
