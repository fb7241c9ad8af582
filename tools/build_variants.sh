#!/bin/bash
# Build A/B variants of the CUDA library with different -D knobs into libcxxdes_b200/_variants/<name>.so:
# the named source files are recompiled with the flags, every other object comes from the regular build (_build/).
# usage: tools/build_variants.sh "file1.cu file2.cu" name1="-DX=1 -DY=0" name2="..."     (run `make` in csrc first)
set -e
cd "$(dirname "$0")/../libcxxdes_b200/csrc"
files="$1"; shift
mkdir -p ../_variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  tmp=$(mktemp -d)
  cp ../_build/*.o $tmp/
  for f in $files; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I../../include -I. $flags -c $f -o $tmp/${f%.cu}.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_variants/$name.so $tmp/*.o -ldl
  rm -rf $tmp
  echo "built _variants/$name.so ($flags)"
done
