set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) 2>&1 | tee gpurun_out/r34_all_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r34_smoke.txt
python bench.py > gpurun_out/r34_bench_default.json 2> gpurun_out/r34_bench.err; cat gpurun_out/r34_bench_default.json; tail -2 gpurun_out/r34_bench.err
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/r34_bench_reference.json
