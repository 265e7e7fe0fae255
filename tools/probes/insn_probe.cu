// Instruction-throughput probe for the softmax inner loop candidates (sm_100a): results per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o insn_probe insn_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 512
#define CHAINS 8

template <int OP>
__global__ void probe(long long* out_clk, uint32_t* sink, uint32_t seed) {
  uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { a[c] = seed + threadIdx.x * 7 + c; b[c] = 0x3c003c00u + c; }
  float fa[CHAINS], fb[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { fa[c] = -0.001f * (threadIdx.x + c); fb[c] = 0.5f + c; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fa[c]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[c]));
      if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[c]));
      if (OP == 3) asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(fa[c]) : "h"((unsigned short)b[c]));
      if (OP == 4) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
      if (OP == 5) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
      if (OP == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(fa[c]) : "f"(fb[c]), "f"(fb[(c + 1) % CHAINS]));
      if (OP == 7) asm volatile("{.reg .b64 t, u; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; add.rn.f32x2 t, t, u; mov.b64 {%0, %1}, t;}" : "+f"(fa[c]), "+f"(fb[c]) : "f"(0.25f));
      if (OP == 8) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[c]) : "f"(__uint_as_float(a[c])), "f"(fb[c]));
      if (OP == 9) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));
      if (OP == 10) asm volatile("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(fa[c]) : "h"((unsigned short)b[c]));
      if (OP == 11) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %1;" : "+f"(fa[c]) : "f"(fb[c]));
      if (OP == 12) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fa[c])); asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(fb[c]) : "h"((unsigned short)b[c])); }
      if (OP == 13) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fa[c])); asm volatile("fma.rn.ftz.f32 %0, %0, %1, %1;" : "+f"(fb[c]) : "f"(0.999f)); }
      if (OP == 14) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[c])); asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c])); }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc ^= a[c] ^ b[c] ^ __float_as_uint(fa[c]) ^ __float_as_uint(fb[c]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_op, long long* clk, uint32_t* sink) {
  const int threads = 1024, blocks = 148;
  long long h[148];
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    probe<OP><<<blocks, threads>>>(clk, sink, rep);
    cudaDeviceSynchronize();
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < blocks; ++i) s += h[i];
    s /= blocks; if (s < best) best = s;
  }
  const double ops = (double)threads * ITERS * CHAINS;
  printf("%-34s %8.1f warp-instr/clk/SM   %7.1f results/clk/SM  (err %s)\n", name, ops / 32.0 / best, ops * per_op / best, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* clk; uint32_t* sink;
  cudaMalloc(&clk, 8 * 1024); cudaMalloc(&sink, 4 * 1024 * 1024);
  run<0>("ex2.approx.ftz.f32", 1, clk, sink);
  run<1>("ex2.approx.f16x2", 2, clk, sink);
  run<2>("ex2.approx.ftz.bf16x2", 2, clk, sink);
  run<3>("add.rn.f32.f16 (FHADD)", 1, clk, sink);
  run<4>("add.rn.f16x2", 2, clk, sink);
  run<5>("max.f16x2", 2, clk, sink);
  run<12>("ex2.f32 + fhadd (pair)", 1, clk, sink);
  run<6>("max.f32 3-input", 2, clk, sink);
  run<7>("add.rn.f32x2", 2, clk, sink);
  run<8>("cvt.rn.f16x2.f32", 2, clk, sink);
  run<9>("fma.rn.f16x2", 2, clk, sink);
  run<10>("fma.rn.f32.f16", 1, clk, sink);
  run<11>("fma.rn.ftz.f32", 1, clk, sink);
  run<13>("ex2.f32 + ffma (pair)", 1, clk, sink);
  run<14>("ex2.f16x2 + hadd2 (pair)", 2, clk, sink);
  return 0;
}
