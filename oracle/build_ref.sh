#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the reference's own CUDA solver (Carabalone/cuda-lbm)
# for its D2Q9 path so that it can be run on the B200 box to (a) generate the golden
# fixtures under tests/golden/ and (b) be timed next to the B200-native solver.
#
# The reference cannot be built as it lies: its 2-D
# path needs the Appendix-B patch set.  This script makes a PRIVATE, TEMPORARY copy of
# /root/reference/src under oracle/_ref/build/ (git-ignored), applies the patches with
# sed, compiles the reference's own translation units + oracle/ref_cuda/ref_driver.cu
# with nvcc for sm_100a, leaves only the binaries in oracle/_ref/bin/ and removes the
# source copy again.  No reference source enters the repository history.
#
# Patches:
#   P1  defines.hpp        replaced by a generated config header (D2Q9, NX/NY/SCALE from -D)
#   P3  boundaries.cuh:11  `float rho` -> `float* rho`               (build fix)
#   P4  zeroGradientOutflow.cuh:13-14  arrays sized [3]              (build fix)
#   P5  lbm_constants.cuh:298  drop the unconditional `#define SOA`  (2-D code is AoS-only)
#   P6  streaming.cuh:9-10 PERIODIC_X/Y selectable with -DREF_PERIODIC_X/-DREF_PERIODIC_Y
#   P8  adapters.cuh:101-107  remove the per-node printf of OptimalAdapter
#   P10 flowPastCylinderScenario.cuh:12  `BGK` -> `BGK<2>`            (build fix; only the s_cyl_* binary includes that file)
# None of them changes the arithmetic of the path.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
BUILD=$OUT/build
BIN=$OUT/bin
if [ ! -d "$REF/src" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
rm -rf "$BUILD"; mkdir -p "$BUILD" "$BIN"
cp -r "$REF/src" "$BUILD/src"
S=$BUILD/src

# P1
cat > "$S/defines.hpp" <<'EOF'
#ifndef DEFINES_H
#define DEFINES_H
#define D2Q9
#ifndef SCALE
#define SCALE 1
#endif
#ifndef NX
#error "NX must be given with -DNX="
#endif
#ifndef NY
#error "NY must be given with -DNY="
#endif
#ifndef NZ
#define NZ 1
#endif
#define BLOCK_SIZE 16
#endif
EOF
# P3
sed -i '11s/float rho, int\* boundary_flags/float* rho, int* boundary_flags/' "$S/core/boundaries/boundaries.cuh"
grep -q 'float\* u, float\* rho, int\* boundary_flags' "$S/core/boundaries/boundaries.cuh"
# P4
sed -i '13s/int normal\[dim\]/int normal[3]/; 14s/int interior_node_coord\[dim\]/int interior_node_coord[3]/' "$S/functors/boundaryConditions/zeroGradientOutflow.cuh"
# P5
sed -i '298s|^#define SOA|// #define SOA (P5)|' "$S/core/lbm_constants.cuh"
! grep -q '^#define SOA' "$S/core/lbm_constants.cuh"
# P6
sed -i '9s|^// #define PERIODIC_X|#ifdef REF_PERIODIC_X\n#define PERIODIC_X\n#endif|; 10s|^// #define PERIODIC_Y|#ifdef REF_PERIODIC_Y\n#define PERIODIC_Y\n#endif|' "$S/core/streaming/streaming.cuh"
grep -q 'REF_PERIODIC_Y' "$S/core/streaming/streaming.cuh"
# P10 (build fix only): flowPastCylinderScenario.cuh:12 names the collision template without its argument
sed -i '12s/BGK/BGK<2>/' "$S/scenarios/flowPastCylinder/flowPastCylinderScenario.cuh"
grep -q 'BGK<2>' "$S/scenarios/flowPastCylinder/flowPastCylinderScenario.cuh"
# P8
sed -i '101,107d' "$S/core/collision/adapters.cuh"
! grep -q 'new_tau_star' "$S/core/collision/adapters.cuh"

NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -rdc=true --std=c++17 -I$S -I$REF/third_party -w"
TUS="core/lbm.cu core/streaming/streaming.cu core/macroscopics/macroscopics.cu core/equilibrium/equilibrium.cu \
     core/collision/MRT/MRT.cu core/collision/CM/CM.cu IBM/IBM_impl.cu IBM/IBM_generators.cu util/utility.cu"

# build_one <name> <defs...>
build_one() {
    local name=$1; shift
    local defs="$*"
    local odir=$BUILD/obj_$name
    mkdir -p "$odir"
    local objs=""
    for tu in $TUS; do
        local o=$odir/$(echo "$tu" | tr '/' '_').o
        $NVCC $FLAGS $defs -c "$S/$tu" -o "$o" &
        objs="$objs $o"
    done
    $NVCC $FLAGS $defs -c "$HERE/ref_cuda/ref_driver.cu" -o "$odir/driver.o" &
    wait
    $NVCC $FLAGS $objs "$odir/driver.o" -o "$BIN/$name" -lcudart
    echo "built $BIN/$name"
}

# The configuration list lives in oracle/ref_cuda/configs.txt: "<name> <defs...>" per line.
while read -r name defs; do
    case "$name" in ''|\#*) continue;; esac
    if [ -n "${ONLY:-}" ] && [[ "$name" != $ONLY ]]; then continue; fi
    build_one "$name" $defs
done < "$HERE/ref_cuda/configs.txt"

rm -rf "$BUILD"
ls -la "$BIN"
