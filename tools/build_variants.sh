#!/bin/bash
# builds liblrce_b200 variants with -D<MACRO>=<n> for same-box timing experiments: tools/build_variants.sh MF_VARIANT 1 2 3
# -> vqa-lrce-kbs-2023_b200/variants/liblrce_<MACRO>_<n>.so (git-ignored; loaded by tools/ through LRCE_LIB=<path>)
set -e
cd "$(dirname "$0")/../vqa-lrce-kbs-2023_b200/csrc"
macro=$1; shift
mkdir -p ../variants
for n in "$@"; do
  rm -rf build_var && mkdir build_var
  for f in host_common gemm_tc mlp_fused rowops window_attn encoder encoder_walk seqops; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -D${macro}=${n} -c $f.cu -o build_var/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/liblrce_${macro}_${n}.so build_var/*.o
  rm -rf build_var
done
