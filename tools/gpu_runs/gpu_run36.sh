set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -12 ) 2>&1 | tee gpurun_out/r36_mp_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r36_bench_n2.json 2> gpurun_out/r36_bench_n2.err; cat gpurun_out/r36_bench_n2.json; tail -3 gpurun_out/r36_bench_n2.err
