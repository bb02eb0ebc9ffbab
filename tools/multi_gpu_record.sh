#!/usr/bin/env bash
# A box with 8 GPUs: strong / weak scaling of the bench workload with the e2e timeline of every rank, PCIe with all GPUs copying at once, the
# unchanged C++ driver on 2 / 4 / 8 GPUs (LBM_B200_GPUS).   gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_record.sh'   -> profiles/r02_multi_gpu.md
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR() { n=$1; port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
# 1. strong scaling of the bench workload at N = 8 with the e2e timeline of every rank
LBM_B200_PIPELINE_DEBUG=1 TR 8 29501 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/multi_n8_bench.json 2> gpurun_out/multi_n8_bench.err
grep -E "lbm_run_from_host" gpurun_out/multi_n8_bench.err | head -8
# the same e2e leg as three plain calls (no band pipeline)
LBM_B200_PIPELINE=0 TR 8 29502 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/multi_n8_bench_threecalls.json 2>> gpurun_out/multi_n8_bench.err
# 2. weak scaling: 4096 rows per GPU
for n in 8 4 2; do TR $n $((29510 + n)) bench.py --gpus $n --steps 40 --warmup 5 --rows-per-gpu 4096 --no-cpu > gpurun_out/multi_n${n}_bench_weak.json 2>> gpurun_out/multi_n8_bench.err; done
# 3. what the ranks share: PCIe with all GPUs copying at once
TR 8 29520 tools/pcie_duplex.py 2>&1 | grep "GPU(s)" | tee gpurun_out/multi_n8_pcie.txt
# 4. the unchanged C++ driver on 8 GPUs (LBM_B200_GPUS): config 4, config 3 (lagged / exact sums, device-side all-reduce), config 5
cd /tmp
( LBM_B200_GPUS=8 timeout 900 $GRAFT_REPO_ROOT/examples/_bin/ex_c4_tg_32768 --steps 40 --warmup 8 --save-int 40 --fast 2>&1 | grep -E "SHIM_RESULT|failed|error," ) 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/multi_n8_shim_c4.txt
for g in 2 4 8; do ( LBM_B200_ADAPTER=1 LBM_B200_GPUS=$g timeout 300 $GRAFT_REPO_ROOT/examples/_bin/ex_c3_lid_4096 --steps 17 --warmup 18 --save-int 17 --repeat 6 --fast 2>&1 | grep -E "SHIM_RESULT|failed" ); done 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/multi_n8_shim_c3_lagged.txt
for g in 2 8; do ( LBM_B200_GPUS=$g timeout 300 $GRAFT_REPO_ROOT/examples/_bin/ex_c3_lid_4096 --steps 17 --warmup 18 --save-int 17 --repeat 6 --fast 2>&1 | grep -E "SHIM_RESULT|failed" ); done 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/multi_n8_shim_c3_exact.txt
( LBM_B200_GPUS=8 timeout 300 $GRAFT_REPO_ROOT/examples/_bin/ex_c5_cyl_8192x2048 --steps 400 --warmup 32 --save-int 400 --fast 2>&1 | grep -E "SHIM_RESULT|failed" ) 2>&1 | tee $GRAFT_REPO_ROOT/gpurun_out/multi_n8_shim_c5.txt
cd $GRAFT_REPO_ROOT
ls -la gpurun_out | tail -14
