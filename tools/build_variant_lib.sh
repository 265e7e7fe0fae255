#!/bin/bash
# experiment builds of the library: tools/build_variant_lib.sh <suffix> <nvcc -D flags...> -> speech-enhancement_b200/libseb200_<suffix>.so (attention_tc.cu recompiled with the flags)
set -e
cd "$(dirname "$0")/.."
sfx=$1; shift
mkdir -p /tmp/t6_$sfx
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c speech-enhancement_b200/csrc/attention_tc.cu -o /tmp/t6_$sfx/attention_tc.o
cd speech-enhancement_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../libseb200_$sfx.so $(ls *.o | grep -v '^attention_tc.o$') /tmp/t6_$sfx/attention_tc.o -lcudart
