#!/bin/bash
# Tuning builds: libcadl_<tag>.so under csrc/variants/ (git-ignored), selected at run time with CADL_LIB=<path>.
# Only one translation unit is recompiled (cadl_stream3.cu, or S2SRC=cadl_api.cu for the other kernels); the other
# objects are shared with the product build.  Knobs: -DCADL_S3_THREADS / _MINB / _DEPTH (CTA shape and ring depth of
# the gradient pass), -DCADL_S3_TRACE (per-warp timeline for profiles/r02_trace.py), -DCADL_PYR_MINB / -DCADL_COEF_MINB,
# -DCADL_PLAN_M / _N / -DCADL_PLAN2_PYR_N (grid splits of the first stage), -DCADL_PT_MINB / _NB (config 2).
# usage: profiles/build_variants.sh "trace:-DCADL_S3_TRACE" "d4:-DCADL_S3_DEPTH=4" ...
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
