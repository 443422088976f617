#!/bin/bash
# Builds libspinnerf_b200.so (sm_100a only) in-tree next to this script's package.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${SPN_LIB_OUT:-$HERE/../libspinnerf_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH=(-gencode arch=compute_100a,code=sm_100a)
FLAGS=(-O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v ${SPN_EXTRA_NVCC_FLAGS:-})
SRCS=(api ops_render mlp_fp32 mlp_tc mlp_tc_bwd peer_reduce)
mkdir -p "$HERE/obj"
pids=()
for f in "${SRCS[@]}"; do
  ( "$NVCC" "${ARCH[@]}" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" > "$HERE/obj/$f.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$HERE"/obj/*.log; exit 1; fi
# the logs are tracked (ptxas -v: registers, spills, shared memory per kernel): drop the only lines that differ from build to build
for f in "${SRCS[@]}"; do sed -i '/Compile time = /d' "$HERE/obj/$f.log"; done
OBJS=()
for f in "${SRCS[@]}"; do OBJS+=("$HERE/obj/$f.o"); done
"$NVCC" "${ARCH[@]}" --shared -o "$OUT" "${OBJS[@]}"
echo "built $OUT"
