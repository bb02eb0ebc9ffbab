# round-1 session-3 verification: GPU tests, smoke, bench (both arms), launch list + one full ncu capture of the bench kernel
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --ny 16384 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n1_half.json 2>/dev/null; cat gpurun_out/bench_n1_half.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 4 -c 2 -o gpurun_out/prof_bench python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out
