#!/bin/bash
# build_variant.sh NAME "EXTRA NVCC FLAGS" -> build/libwn_NAME.so (kernel experiments; select with WN_B200_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
name=$1; shift
tmp=$(mktemp -d)
for f in gmx-acqueduct_b200/csrc/*.cu; do
  extra=""
  case "$f" in *wn_synth.cu) extra="-fmad=false";; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Igmx-acqueduct_b200/csrc $extra "$@" -c "$f" -o "$tmp/$(basename "$f").o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "build/libwn_$name.so" "$tmp"/*.o -ldl
rm -rf "$tmp"
echo "built build/libwn_$name.so"
