#!/bin/bash
# Build libdiqt_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${DIQT_OUT:-$HERE/../libdiqt_b200.so}"          # DIQT_OUT / DIQT_BUILD_DIR / DIQT_EXTRA_FLAGS: A/B variants (tools/gpu_ab.sh)
BUILD="${DIQT_BUILD_DIR:-$HERE/../../build}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr)
[ -n "${DIQT_PTXAS_V:-}" ] && FLAGS+=(-Xptxas -v)
[ -n "${DIQT_EXTRA_FLAGS:-}" ] && FLAGS+=(${DIQT_EXTRA_FLAGS})
mkdir -p "$BUILD"
OBJS=()
PIDS=()
for f in elementwise ends init_tc attn attn_tc linattn_tc backward wgrad_tc volume conv_simt conv_tc conv_zm api; do
  o="$BUILD/$f.o"
  if [ ! -f "$o" ] || [ "$HERE/$f.cu" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/tc_common.cuh" -nt "$o" ] || [ "$HERE/../../include/diqt.h" -nt "$o" ] || [ -n "${DIQT_PTXAS_V:-}" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$o" &
    PIDS+=($!)
  fi
  OBJS+=("$o")
done
for p in "${PIDS[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${OBJS[@]}" -lcudart   # (the arch on the link line too: no default-arch stub cubin in the library)
echo "built $OUT"
