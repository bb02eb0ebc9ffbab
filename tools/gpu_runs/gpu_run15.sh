set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
for c in 0 1 2 3; do tools/kbench 16384 16384 $c 16 | tail -3; done 2>&1 | tee gpurun_out/kbench_packed_o4.txt
tools/kbench 16384 16384 3 16 0 0 | tail -3 | tee -a gpurun_out/kbench_packed_o4.txt
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1_packed_o4.json 2>/dev/null; cat gpurun_out/bench_n1_packed_o4.json
python tools/config_bench.py c1 c2 c3 c3l c5 c5d 2>&1 | tee gpurun_out/config_bench_b.txt
