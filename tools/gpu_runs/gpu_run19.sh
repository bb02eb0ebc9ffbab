set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_ibm_slabs_gpu.py -q -m gpu 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r19_new_tests.txt
for c in c3l c3 c5 c2; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 42 --csv --log-file gpurun_out/r19_launches_$c.csv python tools/config_bench.py $c --steps 40 > /dev/null 2>&1
done
python bench.py --steps 20 --warmup 3 > gpurun_out/r19_bench_n1.json 2>gpurun_out/r19_bench_n1.err; cat gpurun_out/r19_bench_n1.json
