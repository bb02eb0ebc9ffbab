set -x
mkdir -p gpurun_out
export LBM_B200_PIPELINE_DEBUG=1
run() { python bench.py --steps $1 --warmup 3 --no-cpu 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['value'], d['e2e']['seconds'])"; grep run_from_host gpurun_out/err.txt; }
for nb in 128 256 512; do echo "bands $nb"; LBM_B200_PIPELINE_BANDS=$nb run 20; done
echo "steps 40 default bands"; run 40
( timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "run_from_host" 2>&1 | tail -3 )
