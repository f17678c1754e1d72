#!/bin/bash
# Development helper: libpfx variants with different lane-group sizes in the Gaussian kernels.
# usage: build_variants.sh "R F [tag -Dflags...]" ...   -> pulsarfeatureextractor_b200/libpfx_rR_fF[_tag].so
set -e
cd "$(dirname "$0")/../pulsarfeatureextractor_b200/csrc"
make -j8 >/dev/null
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
for v in "$@"; do
  set -- $v
  R=$1; F=$2; TAG=${3:+_$3}; shift; shift; [ $# -gt 0 ] && shift
  ( nvcc $FL -DPFX_ROUNDS_LANES=$R -DPFX_FINAL_LANES=$F "$@" -Xptxas -v -c pfx_k_gauss.cu -o /tmp/g_${R}_${F}${TAG}.o 2> /tmp/g_${R}_${F}${TAG}.log &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpfx_r${R}_f$F$TAG.so pfx_api.o pfx_k_sinusoid.o pfx_k_hist.o /tmp/g_${R}_${F}${TAG}.o pfx_k_dmfit.o pfx_k_subband.o pfx_k_lyon.o &&
    grep "Used" /tmp/g_${R}_${F}${TAG}.log | tr '\n' ' ' && echo " <- r$R f$F $TAG" ) &
done
wait
