set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for c in 0 1 2 3; do tools/kbench 16384 16384 $c 20; done 2>&1 | tee gpurun_out/kbench_r1b.txt
tools/kbench 16384 16384 0 20 1 2>&1 | tee -a gpurun_out/kbench_r1b.txt
tools/kbench 16384 16384 1 20 1 2>&1 | tee -a gpurun_out/kbench_r1b.txt
tools/kbench 32768 32768 0 20 2>&1 | tee -a gpurun_out/kbench_r1b.txt
tools/kbench 4096 4096 0 40 2>&1 | tee -a gpurun_out/kbench_r1b.txt
tools/kbench 256 256 0 200 2>&1 | tee -a gpurun_out/kbench_r1b.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 4 -c 2 -o gpurun_out/prof_r1b tools/kbench 16384 16384 0 2 > gpurun_out/ncu_r1b.log 2>&1; tail -3 gpurun_out/ncu_r1b.log
