#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]   -> underwaterworld_b200/lib/variants/<name>.so
# name "main" builds the product library underwaterworld_b200/lib/libuwcuda.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
out="$ROOT/underwaterworld_b200/lib/variants/$name.so"
[ "$name" = "main" ] && out="$ROOT/underwaterworld_b200/lib/libuwcuda.so"
mkdir -p "$ROOT/underwaterworld_b200/lib/variants"
cd "$ROOT/underwaterworld_b200/csrc"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -diag-suppress 177 "$@" -o "$out" uwcuda.cu
echo "built $out"
