#!/bin/bash
# Builds tools/ab/lib_kmprof.so: the library with -DKM_PROFILE (clock64 phase counters in the
# k-means kernels; printed by spalign_kmeans_debug_stats).  Use with SPALIGN_LIB=tools/ab/lib_kmprof.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/ab /tmp/kmprof
for s in superpixel_align_b200/csrc/*.cu; do
  f=$(basename $s .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DKM_PROFILE \
    -I include -c $s -o /tmp/kmprof/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/ab/lib_kmprof.so /tmp/kmprof/*.o
