#!/bin/bash
# debug build of the library with the tcgen05 attention kernel's per-tile clock64 trace enabled for one CTA
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/t5
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DT5_TRACE=${1:-20000} -c speech-enhancement_b200/csrc/attention_tc.cu -o /tmp/t5/attention_tc_trace.o
cd speech-enhancement_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../libseb200_trace.so core.o gemm_api.o tok_gemm.o conv_y3.o conv_persist.o conv_tap.o ffn_fused.o dsp.o norm_act.o attention.o /tmp/t5/attention_tc_trace.o dwconv.o dwpw2.o -lcudart
