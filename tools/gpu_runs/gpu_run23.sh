set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/r23_all_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r23_smoke.txt
cd examples/_bin && ./ex_c1_tg_256 --steps 600 --save-int 300 --fast --save-ckpt /tmp/c1.ckpt | grep -E "error|checkpoint|SHIM" ; ./ex_c1_tg_256 --steps 400 --save-int 400 --fast --load-ckpt /tmp/c1.ckpt | grep -E "error|restart|SHIM"; ./ex_c1_tg_256 --steps 1000 --save-int 1000 --fast | grep -E "error|SHIM"
