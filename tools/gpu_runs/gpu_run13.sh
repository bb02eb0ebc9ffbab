# packed-fp32 (FFMA2) collision kernels: parity + per-operator kernel timings + bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
for c in 0 1 2 3; do tools/kbench 16384 16384 $c 16 | tail -3; done 2>&1 | tee gpurun_out/kbench_packed.txt
tools/kbench 16384 16384 3 16 0 0 | tail -3 | tee -a gpurun_out/kbench_packed.txt
tools/kbench 16384 16384 0 16 1 | tail -3 | tee -a gpurun_out/kbench_packed.txt
tools/kbench 16384 16384 1 16 1 | tail -3 | tee -a gpurun_out/kbench_packed.txt
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1_packed.json 2>/dev/null; cat gpurun_out/bench_n1_packed.json
