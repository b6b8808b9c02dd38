#!/bin/bash
# A/B libraries: tools/build_ab.sh NAME SRC.cu [-DFLAG ...] builds tools/ab/lib_NAME.so from the
# standard objects (csrc/build/*.o, built by __graft_entry__.build()) with SRC.cu recompiled with
# the given flags in place of the object of the same base name.  Use with SPALIGN_LIB=tools/ab/lib_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
base=$(basename $src .cu)
mkdir -p tools/ab /tmp/ab_$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" \
  -I include -I superpixel_align_b200/csrc -c $src -o /tmp/ab_$name/$base.o
objs=$(ls superpixel_align_b200/csrc/build/*.o | grep -v "/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/ab/lib_$name.so $objs /tmp/ab_$name/$base.o
