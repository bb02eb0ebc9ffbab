# 2-GPU box: shim tests (GPU 0), then the slab-coupling diagnosis at N=2
set -x
mkdir -p gpurun_out
python -m pytest tests/test_shim_gpu.py -q -s 2>&1 | grep -E "vs |analytic|Poiseuille 64|after 300|passed|failed|Error" | tee gpurun_out/shim_parity.txt
nvidia-smi topo -m | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
$TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/n2_direct.json 2> gpurun_out/n2_direct.err; cat gpurun_out/n2_direct.json; tail -3 gpurun_out/n2_direct.err
LBM_SLAB_MODE=nccl $TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/n2_nccl.json 2> gpurun_out/n2_nccl.err; cat gpurun_out/n2_nccl.json; tail -3 gpurun_out/n2_nccl.err
# two independent half-size single-GPU runs at the same time: what the two GPUs deliver without any coupling
CUDA_VISIBLE_DEVICES=0 python bench.py --ny 16384 --steps 40 --warmup 3 --no-cpu --no-e2e > gpurun_out/indep0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python bench.py --ny 16384 --steps 40 --warmup 3 --no-cpu --no-e2e > gpurun_out/indep1.json 2>/dev/null &
wait
cat gpurun_out/indep0.json gpurun_out/indep1.json
LBM_B200_NO_HANDSHAKE=1 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/n2_nohs.json 2> gpurun_out/n2_nohs.err; cat gpurun_out/n2_nohs.json; tail -3 gpurun_out/n2_nohs.err
