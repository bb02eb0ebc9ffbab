set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
$TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_direct.json 2> gpurun_out/bench_n2_direct.err; cat gpurun_out/bench_n2_direct.json; tail -5 gpurun_out/bench_n2_direct.err
LBM_SLAB_MODE=nccl $TR --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; cat gpurun_out/bench_n2_nccl.json; tail -5 gpurun_out/bench_n2_nccl.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
