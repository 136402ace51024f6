#!/usr/bin/env bash
# Builds libfsg_dense.so (sm_100a only) next to the package.  Usage: csrc/build.sh [extra nvcc flags]
# The translation units are compiled in parallel and linked into one shared library; FSG_OUT overrides the output
# path (kernel experiments build variants side by side and select one with FSG_DENSE_LIB).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${FSG_OUT:-${here}/../libfsg_dense.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=(abi nms_large iou_match dense_loss dense_step bet_stats dense_loss_levels detect_select nms_image rpn_select layout)
objdir="$(mktemp -d "${TMPDIR:-/tmp}/fsg_build.XXXXXX")"
trap 'rm -rf "${objdir}"' EXIT
flags=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --fmad=true "$@")
pids=()
for s in "${srcs[@]}"; do
  "${NVCC}" "${flags[@]}" -c -o "${objdir}/${s}.o" "${here}/${s}.cu" > "${objdir}/${s}.log" 2>&1 &
  pids+=($!)
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then fail=1; fi
  cat "${objdir}/${srcs[$i]}.log"
done
if [ "${fail}" -ne 0 ]; then echo "nvcc failed" >&2; exit 1; fi
objs=()
for s in "${srcs[@]}"; do objs+=("${objdir}/${s}.o"); done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "${out}" "${objs[@]}"
echo "built ${out}"
