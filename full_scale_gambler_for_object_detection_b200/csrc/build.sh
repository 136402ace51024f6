#!/usr/bin/env bash
# Builds libfsg_dense.so (sm_100a only) next to the package.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${FSG_OUT:-${here}/../libfsg_dense.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=("${here}/abi.cu" "${here}/nms_large.cu" "${here}/iou_match.cu" "${here}/dense_loss.cu" "${here}/dense_loss_levels.cu" "${here}/detect_select.cu" "${here}/nms_image.cu" "${here}/rpn_select.cu" "${here}/layout.cu")
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
  --fmad=true "$@" -o "${out}" "${srcs[@]}"
echo "built ${out}"
