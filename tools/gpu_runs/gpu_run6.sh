mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,power.draw,temperature.gpu --format=csv
tools/kbench 16384 16384 0 20
tools/kbench 32768 32768 0 10
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,power.draw,temperature.gpu --format=csv
