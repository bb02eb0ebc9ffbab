#!/usr/bin/env bash
# Development aid: builds the engine with alternative compile-time switches into tools/variants/<name>/liblbm_b200.so
# (git-ignored; the files travel to the GPU box) so that tools/kbench can A/B them with LD_LIBRARY_PATH.
#   tools/build_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags ...]
set -euo pipefail
cd "$(dirname "$0")/.."
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    mkdir -p tools/variants/$name
    ( nvcc $ARCH -O3 -lineinfo --std=c++17 -fmad=false -Xcompiler -fPIC $flags -shared -o tools/variants/$name/liblbm_b200.so cuda_lbm_b200/csrc/engine.cu -lcudart && echo "built $name ($flags)" ) &
done
wait
