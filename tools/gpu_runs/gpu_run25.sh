set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r25_all_tests.txt
LBM_B200_GRAPH=0 python tools/config_bench.py c1 c2 c5 c3l --steps 400 2>&1 | grep '^{' | tee gpurun_out/r25_config_bench_nograph.txt
python tools/config_bench.py c1 c2 c5 c3l c3 c5d --steps 400 2>&1 | grep '^{' | tee gpurun_out/r25_config_bench_graph.txt
LBM_B200_GRAPH=1 python tools/config_bench.py c5 c3l --steps 400 2>&1 | grep '^{' | tee gpurun_out/r25_config_bench_graph_forced.txt
