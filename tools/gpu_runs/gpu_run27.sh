set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r27_all_tests.txt
python tools/config_bench.py c3 c3l c5 --steps 400 2>&1 | grep '^{' | tee gpurun_out/r27_config_bench.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r27_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 4 -c 2 -o gpurun_out/r27_prof_bench python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/
