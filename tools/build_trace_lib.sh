#!/bin/bash
# debug build of the library with the tcgen05 attention kernel's per-tile clock64 trace enabled for one CTA (-DT6_TRACE=<block>)
set -e
cd "$(dirname "$0")/.."
python speech-enhancement_b200/build.py
mkdir -p /tmp/t6
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DT6_TRACE=${1:-25001} -c speech-enhancement_b200/csrc/attention_tc.cu -o /tmp/t6/attention_tc_trace.o
cd speech-enhancement_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../libseb200_trace.so $(ls *.o | grep -v '^attention_tc.o$') /tmp/t6/attention_tc_trace.o -lcudart
