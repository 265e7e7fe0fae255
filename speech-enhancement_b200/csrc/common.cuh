// common.cuh -- shared helpers for libseb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/seb200.h"

namespace seb {

// ---- error plumbing (never abort, never print: seb200.h conventions) ----------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SEB_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      seb::set_error(__VA_ARGS__);              \
      return (code);                            \
    }                                           \
  } while (0)

#define SEB_CHECK_LAUNCH(name)                                                  \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      seb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return (int)e__;                                                          \
    }                                                                           \
    seb::count_launch();                                                        \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- small device helpers -------------------------------------------------------
// sigmoid via ex2.approx + rcp.approx (2 MUFU + 2 FMA-pipe ops; ~2 ulp, far inside the 1e-3 path tolerance)
__device__ __forceinline__ float sigmoidf_acc(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// tanh(x) = 1 - 2 / (1 + e^(2x)) on the same two MUFU ops (absolute error ~2e-7: the result multiplies a sigmoid in (0, 1))
__device__ __forceinline__ float tanhf_acc(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(2.8853900817779268f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fp32 -> (hi, lo) bf16 pair with hi + lo ~= x to ~2^-17 relative (round-to-nearest both)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// fp32 -> (hi, mid, lo) bf16 triple: hi + mid + lo == x to ~2^-24 relative
__device__ __forceinline__ void split3_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  float2 hf = __bfloat1622float2(h);
  const float r0 = x0 - hf.x, r1 = x1 - hf.y;
  __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
  float2 mf = __bfloat1622float2(m);
  __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  mid = *reinterpret_cast<uint32_t*>(&m);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace seb
