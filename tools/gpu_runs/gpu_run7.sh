mkdir -p gpurun_out
echo "== default (e6 o5)"; for c in 0 1 2 3; do tools/kbench 16384 16384 $c 16 | tail -3; done
for v in e6o4 e6o6 e5o5 e4o4; do echo "== $v"; for c in 0 2 3; do LD_LIBRARY_PATH=tools/variants/$v tools/kbench 16384 16384 $c 16 | tail -3; done; done
