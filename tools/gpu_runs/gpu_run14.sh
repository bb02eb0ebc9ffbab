set -x
mkdir -p gpurun_out
for v in base o4 o6 e5o4 ldcs stcs ldcs_stcs ldlu ldlu_stcs stwt; do echo "== $v"; for c in 0 3; do LD_LIBRARY_PATH=tools/variants/$v tools/kbench 16384 16384 $c 16 | tail -3; done; done 2>&1 | tee gpurun_out/kbench_variants_packed.txt
python -m pytest tests/test_shim_gpu.py -q -x -k optimal_adapter -s 2>&1 | tail -5
