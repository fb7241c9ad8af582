#!/bin/bash
# Build A/B variants of the CUDA library with different -D knobs into libcxxdes_b200/_variants/<name>.so
# usage: tools/build_variants.sh name1="-DX=1 -DY=0" name2="..."
set -e
cd "$(dirname "$0")/../libcxxdes_b200/csrc"
mkdir -p ../_variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  rm -rf /tmp/cxb_variant_build && mkdir -p /tmp/cxb_variant_build
  for f in c_api mm1_fused mm1_ahead mm1_engine reduce program_kernel models_aloha models_archsim models_mmc; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I../../include -I. $flags -c $f.cu -o /tmp/cxb_variant_build/$f.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_variants/$name.so /tmp/cxb_variant_build/*.o
  echo "built _variants/$name.so ($flags)"
done
