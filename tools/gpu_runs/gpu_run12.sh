# 8-GPU box: the contract bench at N=8 (direct peer-mapped coupling) and N=4
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 40 --warmup 4 > gpurun_out/n8_direct.json 2> gpurun_out/n8_direct.err; cat gpurun_out/n8_direct.json; tail -5 gpurun_out/n8_direct.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 40 --warmup 4 --no-e2e > gpurun_out/n4_direct.json 2> gpurun_out/n4_direct.err; cat gpurun_out/n4_direct.json; tail -5 gpurun_out/n4_direct.err
