set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu -k "run_from_host or c2_pois" 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r29_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r29_bench_n1.json 2>gpurun_out/r29_bench.err; cat gpurun_out/r29_bench_n1.json; tail -3 gpurun_out/r29_bench.err
for nb in 128 256 1024; do LBM_B200_PIPELINE_BANDS=$nb python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bands $nb', d['e2e'])"; done | tee gpurun_out/r29_bands.txt
LBM_B200_PIPELINE=0 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline off', d['e2e'])" | tee -a gpurun_out/r29_bands.txt
python bench.py --steps 40 --warmup 4 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('steps 40', d['value'], d['e2e'])" | tee -a gpurun_out/r29_bands.txt
