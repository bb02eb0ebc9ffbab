set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/r39.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e'])" | tee gpurun_out/r39_bench.txt; tail -2 gpurun_out/r39.err
( timeout 120 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "peer_handshake or run_from_host_on_peer or engine_matches_oracle" 2>&1 | tail -3 ) | tee gpurun_out/r39_tests.txt
