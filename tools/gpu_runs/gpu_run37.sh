set -x
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "run_from_host" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r37_tests.txt
python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds'])" | tee gpurun_out/r37_bench.txt
