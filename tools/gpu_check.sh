#!/usr/bin/env bash
# One GPU: smoke, the whole GPU test suite, the default bench line and the reference arm (what the driver runs at round end).
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'        outputs under gpurun_out/
set -x
mkdir -p gpurun_out
rm -f gpurun_out/fullsize_parity.txt
python __graft_entry__.py smoke 2>&1 | tail -6 | tee gpurun_out/check_smoke.txt
( timeout 1700 python -m pytest tests -m gpu -q --no-header -rf --durations=5 2>&1 | tail -15 ) > gpurun_out/check_tests.txt 2>&1
tail -3 gpurun_out/check_tests.txt
timeout 900 python bench.py > gpurun_out/check_bench_default.json 2> gpurun_out/check_bench.err; tail -2 gpurun_out/check_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/check_bench_ref_default.json 2>> gpurun_out/check_bench.err
