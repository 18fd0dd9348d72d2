#!/bin/sh
# Stage the reference's own sm100 HSTU kernels (pure-Python CuTe-DSL, 14 files) where the GPU box can import them:
# baseline/_ref/ is git-ignored but travels with gpurun.  Nothing here is product source; used only by tools/bench_vs_reference.py.
set -e
SRC=/root/reference/third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell
DST="$(dirname "$0")/_ref/hstu_blackwell"
mkdir -p "$(dirname "$DST")"
rm -rf "$DST"
cp -r "$SRC" "$DST"
echo "staged $(ls "$DST" | wc -l) files in $DST"
