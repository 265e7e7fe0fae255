#!/bin/bash
# experiment builds of the library: tools/build_variant_lib.sh <suffix> <source.cu> <nvcc -D flags...> -> speech-enhancement_b200/libseb200_<suffix>.so
# (the named source recompiled with the flags, every other object taken from the regular build); SEB200_LIB_SUFFIX=<suffix> selects it at import
set -e
cd "$(dirname "$0")/.."
sfx=$1; src=$2; shift 2
python speech-enhancement_b200/build.py > /dev/null
mkdir -p /tmp/seb_$sfx
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c speech-enhancement_b200/csrc/$src -o /tmp/seb_$sfx/${src%.cu}.o
cd speech-enhancement_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../libseb200_$sfx.so $(ls *.o | grep -v "^${src%.cu}.o$") /tmp/seb_$sfx/${src%.cu}.o -lcudart
