set -x
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -30 ) 2>&1 | tee gpurun_out/r22_mp_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mp_slab_worker.py --kind ibm --mode direct --coll 1 2>&1 | grep MP_PARITY | tee gpurun_out/r22_mp_line.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r22_bench_n2.json 2> gpurun_out/r22_bench_n2.err; cat gpurun_out/r22_bench_n2.json
