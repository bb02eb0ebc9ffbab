set -x
mkdir -p gpurun_out
export LBM_B200_PIPELINE_DEBUG=1
run() { python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 >/dev/null | grep run_from_host; }
echo "default"; run
echo "nocopy"; LBM_B200_PIPELINE_NOCOPY=1 run
echo "noskew"; LBM_B200_PIPELINE_NOSKEW=1 run
echo "noskew nocopy"; LBM_B200_PIPELINE_NOSKEW=1 LBM_B200_PIPELINE_NOCOPY=1 run
echo "noskew nocopy 64 bands"; LBM_B200_PIPELINE_BANDS=64 LBM_B200_PIPELINE_NOSKEW=1 LBM_B200_PIPELINE_NOCOPY=1 run
echo "nocopy 64 bands"; LBM_B200_PIPELINE_BANDS=64 LBM_B200_PIPELINE_NOCOPY=1 run
