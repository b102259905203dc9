#!/bin/bash
# Tuning builds of the streaming gradient kernel: libcadl_<tag>.so under csrc/variants/ (git-ignored), selected at run
# time with CADL_LIB=<path>.  Only cadl_stream2.cu is recompiled; the other objects are shared with the product build.
# usage: profiles/build_variants.sh "tag:-DCADL_S2_MINB=3" "tag2:-DCADL_S2_PF=8 ..." ...
set -e
CS="$(dirname "$0")/../camera-aware-neural-networks-for-few-view-depth-estimation_b200/csrc"
cd "$CS"
make -j4 >/dev/null
mkdir -p variants
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  src="${S2SRC:-cadl_stream3.cu}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include $flags \
       -c -o variants/${src%.cu}_$tag.o $src -Xptxas -v 2> variants/$tag.ptxas.log
  objs=$(ls obj/*.o | grep -v "${src%.cu}.o" | grep -v "_dbg.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libcadl_$tag.so $objs variants/${src%.cu}_$tag.o
  echo "built variants/libcadl_$tag.so ($flags)"
done
