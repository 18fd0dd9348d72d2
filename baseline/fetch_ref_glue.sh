#!/bin/sh
# Stage the reference's own Triton glue kernels of the HSTU layer (layer norm, LN*u*dropout; pure Python/Triton, 3 files) where the GPU box
# can import them: baseline/_ref/ is git-ignored but travels with gpurun.  Nothing here is product source; used only by
# tools/e2e_harness.py's "reference_fused" arm (the reference's fused layer restated on the reference's own kernels).
set -e
EX=/root/reference/examples
DST="$(dirname "$0")/_ref/refglue"
rm -rf "$DST"
mkdir -p "$DST/commons/ops/triton_ops" "$DST/ops/triton_ops"
cp "$EX/commons/ops/triton_ops/common.py" "$DST/commons/ops/triton_ops/"
cp "$EX/hstu/ops/triton_ops/triton_layer_norm.py" "$EX/hstu/ops/triton_ops/triton_norm_mul_dropout.py" "$DST/ops/triton_ops/"
echo "staged $(find "$DST" -name '*.py' | wc -l) files in $DST"
