// common.cuh -- shared helpers for libseb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

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

// Once-per-(call site, device) guard.  Kernel attributes (opt-in dynamic shared memory) are per device and one process may drive
// several devices (one module instance per cuda:k), so a plain `static bool` would leave every device but
// the first without its attribute.  Racing threads at worst set the same attribute twice.
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool done() const {
    int d = 0;
    return cudaGetDevice(&d) == cudaSuccess && d < 64 && ((mask.load(std::memory_order_acquire) >> d) & 1ull);
  }
  void set() {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && d < 64) mask.fetch_or(1ull << d, std::memory_order_release);
  }
};

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

// fp32 -> (hi, lo) bf16 pair with hi + lo ~= x to ~2^-17 relative (round-to-nearest both).  The residual x - hi is exact in
// fp32 either way; it is formed with one packed FFMA2 (hi * -1 + x) for the two values instead of two FADDs.
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float2 hf = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  const float2 r = __ffma2_rn(hf, make_float2(-1.0f, -1.0f), make_float2(x0, x1));
  __nv_bfloat162 l = __floats2bfloat162_rn(r.x, r.y);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// LayerNorm of four values on packed fp32x2: ((v - mean) * rstd) * gamma + beta in the reference's operation order
__device__ __forceinline__ void ln_apply4(const float4 v, float mean, float rstd, const float4 g, const float4 b, float2& y01, float2& y23) {
  const float2 nm = make_float2(-mean, -mean), rs = make_float2(rstd, rstd);
  y01 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v.x, v.y), nm), rs), make_float2(g.x, g.y), make_float2(b.x, b.y));
  y23 = __ffma2_rn(__fmul2_rn(__fadd2_rn(make_float2(v.z, v.w), nm), rs), make_float2(g.z, g.w), make_float2(b.z, b.w));
}

// shifted one-pass statistics of four more values: s += (v - x0), q += (v - x0)^2, both as float2 partial sums
__device__ __forceinline__ void stats_acc4(const float4 v, float2 nx0, float2& s, float2& q) {
  const float2 d0 = __fadd2_rn(make_float2(v.x, v.y), nx0), d1 = __fadd2_rn(make_float2(v.z, v.w), nx0);
  s = __fadd2_rn(s, __fadd2_rn(d0, d1));
  q = __ffma2_rn(d0, d0, q);
  q = __ffma2_rn(d1, d1, q);
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
