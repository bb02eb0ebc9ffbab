#!/usr/bin/env bash
# Runs the reference's own CUDA solver (oracle/_ref/bin/g_*, built by oracle/build_ref.sh from the sources
# under /root/reference) on the GPU of this box and leaves raw dumps in gpurun_out/golden/.
# tests/golden/pack_golden.py then turns them into the committed tests/golden/<case>.npz fixtures.
#   usage (from the repo root, under gpurun):  bash tests/golden/make_golden.sh
set -uo pipefail
OUT=gpurun_out/golden
mkdir -p "$OUT"
for b in oracle/_ref/bin/g_*; do
    name=$(basename "$b")
    steps=100; dumps="0 1 2 3 10 100"
    case "$name" in *cmopt*) steps=30; dumps="0 1 2 3 10 30";; esac
    echo "== $name"
    timeout 300 "$b" $steps "$OUT" "$name" $dumps > "$OUT/$name.log" 2>&1 || echo "FAILED $name"
    grep REF_MLUPS "$OUT/$name.log" || tail -3 "$OUT/$name.log"
done
ls "$OUT" | wc -l
