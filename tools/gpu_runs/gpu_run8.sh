mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
tools/kbench 16384 16384 3 16 | tail -3
tools/kbench 16384 16384 3 16 0 0 | tail -3
python tools/config_bench.py 2>&1 | tee gpurun_out/config_bench_r1a.txt
