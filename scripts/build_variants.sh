#!/bin/bash
# Development helper: libpfx variants with one translation unit rebuilt with extra -D flags.
# usage: build_variants.sh "unit tag -Dflags..." ...   e.g. "pfx_k_gauss r2 -DPFX_ROUNDS_LANES=2"
#   -> pulsarfeatureextractor_b200/libpfx_<tag>.so
set -e
cd "$(dirname "$0")/../pulsarfeatureextractor_b200/csrc"
make -j8 >/dev/null
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
ALL="pfx_api pfx_k_sinusoid pfx_k_hist pfx_k_gauss pfx_k_dmfit pfx_k_dmfit_exact pfx_k_subband pfx_k_lyon"
for v in "$@"; do
  set -- $v
  U=$1; TAG=$2; shift; shift
  OBJS=""
  for o in $ALL; do if [ $o = $U ]; then OBJS="$OBJS /tmp/${U}_${TAG}.o"; else OBJS="$OBJS $o.o"; fi; done
  ( nvcc $FL "$@" -Xptxas -v -c $U.cu -o /tmp/${U}_${TAG}.o 2> /tmp/${U}_${TAG}.log &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpfx_${TAG}.so $OBJS &&
    grep "Used" /tmp/${U}_${TAG}.log | tr '\n' ' ' && echo " <- $U $TAG" ) &
done
wait
