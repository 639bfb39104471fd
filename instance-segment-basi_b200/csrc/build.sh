#!/bin/bash
# Builds libbasi_b200.so (sm_100a) in-tree.  Usage: csrc/build.sh [extra nvcc flags]
# Every .cu is compiled to an object in parallel (objects newer than their source and the headers are reused), then
# linked into one shared library.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../basi_b200/libbasi_b200.so"
OBJ="$HERE/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -diag-suppress 177,550 -I"$ROOT/include" -I"$HERE" "$@")
mkdir -p "$OBJ"
echo "${FLAGS[*]}" > "$OBJ/flags.new"
if ! cmp -s "$OBJ/flags.new" "$OBJ/flags"; then rm -f "$OBJ"/*.o; mv "$OBJ/flags.new" "$OBJ/flags"; fi
pids=()
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
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static "$OBJ"/*.o -o "$OUT"
echo "built $OUT"
