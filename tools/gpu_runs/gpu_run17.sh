set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_v2.json 2>/dev/null; cat gpurun_out/bench_n1_v2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_bench_v2.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 4 -c 2 -o gpurun_out/prof_bench_v2 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_c5.csv python tools/config_bench.py c5d --steps 20 > /dev/null 2>&1
for c in 0 1 2 3; do tools/kbench 16384 16384 $c 16 | tail -3; done 2>&1 | tee gpurun_out/kbench_v2.txt
tools/kbench 16384 16384 3 16 0 0 | tail -3 | tee -a gpurun_out/kbench_v2.txt
