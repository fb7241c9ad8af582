#!/bin/bash
# Is the device code of the working tree the code that was measured?  Compiles every translation unit of
# libcxxdes_b200/csrc at a git revision and in the working tree for sm_100a and compares the SASS (addresses and the
# hash nvcc puts into the names of anonymous namespaces removed).  No GPU needed.
#   tools/sass_same.sh <git revision>          e.g. the commit whose numbers are in profiles/
set -e
cd "$(dirname "$0")/.."
rev=${1:?usage: tools/sass_same.sh <git revision>}
old=$(mktemp -d)
git archive "$rev" libcxxdes_b200/csrc include | tar -x -C "$old"
flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC"
sass() { cuobjdump -sass "$1" | sed 's#/\*[0-9a-f]*\*/##g; s/_GLOBAL__N__[0-9a-f]*_[0-9]*_[a-z_0-9]*_cu_[0-9a-f]*//g' | grep -v '^\s*$' | md5sum | cut -c1-16; }
for tree in "$old" "$PWD"; do
    mkdir -p "$old/obj$([ "$tree" = "$old" ] && echo _old || echo _new)"
done
for cu in libcxxdes_b200/csrc/*.cu; do
    f=$(basename "$cu" .cu)
    [ "$f" = c_api ] && continue        # host code only
    ( [ -f "$old/$cu" ] && cd "$old/libcxxdes_b200/csrc" && nvcc $flags -I../../include -I. -c "$f.cu" -o "$old/obj_old/$f.o" 2>/dev/null ) &
    ( cd libcxxdes_b200/csrc && nvcc $flags -I../../include -I. -c "$f.cu" -o "$old/obj_new/$f.o" 2>/dev/null ) &
done
wait
for cu in libcxxdes_b200/csrc/*.cu; do
    f=$(basename "$cu" .cu)
    [ "$f" = c_api ] && continue
    if [ ! -f "$old/obj_old/$f.o" ]; then echo "$f: new file"; continue; fi
    a=$(sass "$old/obj_old/$f.o"); b=$(sass "$old/obj_new/$f.o")
    echo "$f: $([ "$a" = "$b" ] && echo same || echo DIFFERENT)"
done
rm -rf "$old"
