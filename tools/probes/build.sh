#!/bin/bash
# builds the stand-alone probes next to their sources (sm_100a only; binaries are git-ignored but travel with gpurun)
set -e
cd "$(dirname "$0")"
CS=../../vqa-lrce-kbs-2023_b200/csrc
for p in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -o ${p}.bin ${p}.cu $CS/host_common.cu -lcuda
done
