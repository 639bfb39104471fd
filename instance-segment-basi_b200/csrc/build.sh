#!/bin/bash
# Builds the in-tree libraries (sm_100a).  Usage: csrc/build.sh [extra nvcc flags]
#   basi_b200/libbasi_b200.so      16-bit storage type = bfloat16 (default; also hosts the split-operand f32 mode)
#   basi_b200/libbasi_b200_f16.so  the same sources with -DBASI_HALF_FP16: 16-bit storage type = IEEE fp16
# Every .cu is compiled to an object in parallel (objects newer than their source and the headers are reused), then
# linked into one shared library per format.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
BASEFLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
           -diag-suppress 177,550 -I"$ROOT/include" -I"$HERE" "$@")
pids=()
for variant in bf16 f16; do
  if [ "$variant" = f16 ]; then OBJ="$HERE/build/f16"; FLAGS=("${BASEFLAGS[@]}" -DBASI_HALF_FP16); else OBJ="$HERE/build/bf16"; FLAGS=("${BASEFLAGS[@]}"); fi
  mkdir -p "$OBJ"
  echo "${FLAGS[*]}" > "$OBJ/flags.new"
  if ! cmp -s "$OBJ/flags.new" "$OBJ/flags"; then rm -f "$OBJ"/*.o; mv "$OBJ/flags.new" "$OBJ/flags"; fi
  for src in "$HERE"/*.cu; do
    obj="$OBJ/$(basename "${src%.cu}").o"
    stale=0
    [ -f "$obj" ] || stale=1
    for dep in "$src" "$HERE"/*.cuh "$ROOT"/include/*.h; do
      [ "$stale" = 1 ] || { [ "$dep" -nt "$obj" ] && stale=1; } || true
    done
    if [ "$stale" = 1 ]; then
      "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
      pids+=($!)
    fi
  done
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static "$HERE"/build/bf16/*.o -o "$HERE/../basi_b200/libbasi_b200.so"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static "$HERE"/build/f16/*.o -o "$HERE/../basi_b200/libbasi_b200_f16.so"
echo "built $HERE/../basi_b200/libbasi_b200.so and libbasi_b200_f16.so"
