set -x
for i in 1 2; do
for c in 0 2; do echo "== current coll $c"; tools/kbench 16384 16384 $c 16 | tail -3; echo "== prerefactor coll $c"; LD_LIBRARY_PATH=tools/variants/o4_prerefactor tools/kbench 16384 16384 $c 16 | tail -3; done
done 2>&1 | tee gpurun_out/kbench_ab_refactor.txt
