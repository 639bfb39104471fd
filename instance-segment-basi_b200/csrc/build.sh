#!/bin/bash
# Builds libbasi_b200.so (sm_100a) in-tree.  Usage: csrc/build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../basi_b200/libbasi_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=("$HERE"/*.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -cudart static \
  -I"$ROOT/include" -I"$HERE" "$@" "${SRCS[@]}" -o "$OUT"
echo "built $OUT"
