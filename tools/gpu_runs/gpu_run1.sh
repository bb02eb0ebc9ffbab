set -x
nvidia-smi -L
bash tests/golden/make_golden.sh 2>&1 | tail -40
python tests/gpu_report.py > gpurun_out/gpu_report.txt 2>&1; tail -40 gpurun_out/gpu_report.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --nx 8192 --ny 8192 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_8192.json 2> gpurun_out/bench_8192.err; cat gpurun_out/bench_8192.json; tail -3 gpurun_out/bench_8192.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
for b in c1_tg_bgk_256 c2_pois_mrt_1024x256 c3_lid_cmopt_4096 c5_cyl_ibm_mrt_8192x2048 t_tg_bgk_4096 t_tg_bgk_8192; do timeout 300 oracle/_ref/bin/$b 60 /tmp x | grep REF_MLUPS; done > gpurun_out/ref_mlups.txt 2>&1; cat gpurun_out/ref_mlups.txt
