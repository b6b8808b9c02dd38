#!/bin/bash
# Builds tools/ab/lib_kmprof.so: the library with -DKM_PROFILE (clock64 phase counters in the
# k-means kernels; printed by spalign_kmeans_debug_stats).  Use with SPALIGN_LIB=tools/ab/lib_kmprof.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/ab /tmp/kmprof
for f in overlap pool paint kmeans; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DKM_PROFILE \
    -I include -c superpixel_align_b200/csrc/$f.cu -o /tmp/kmprof/$f.o &
done
wait
nvcc -shared -o tools/ab/lib_kmprof.so /tmp/kmprof/*.o
