set -x
mkdir -p gpurun_out
python tools/pcie_duplex.py 2>&1 | tee gpurun_out/r30_pcie.txt
for nb in 256 512; do LBM_B200_PIPELINE_DEBUG=1 LBM_B200_PIPELINE_BANDS=$nb python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 >/dev/null | grep run_from_host; done | tee gpurun_out/r30_pipeline_debug.txt
LBM_B200_PIPELINE_DEBUG=1 python bench.py --steps 40 --warmup 3 --no-cpu 2>&1 >/dev/null | grep run_from_host | tee -a gpurun_out/r30_pipeline_debug.txt
( timeout 600 python -m pytest tests/test_full_size_gpu.py -q -m gpu -k "c2_pois" 2>&1 | tail -3 ) | tee gpurun_out/r30_c2.txt
