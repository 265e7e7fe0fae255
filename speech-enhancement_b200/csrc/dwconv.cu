// dwconv.cu -- DepthWiseConv1d(128, k=31, zero pad 15/15) + BatchNorm1d(eval, folded) + Swish along the sequence
// axis (conformer.py:166-168).  One CTA = 64 positions x 128 channels of one sequence; thread = channel; the
// (64+30) x 128 input tile is staged in shared memory (coalesced 512-byte rows), each thread produces 8 outputs per
// pass from a 38-row register window with packed fp32x2 FMAs (two channels per instruction).
#include "common.cuh"
#include <type_traits>

namespace seb {

constexpr int DW_TI = 64, DW_K = 31, DW_PAD = 15, DW_C = 128;

// Thread = (channel pair, half of the 64 positions).  All arithmetic is packed fp32x2 (FFMA2, sm_100): two channels per
// instruction.  Per group of 8 outputs the 38-row input window is consumed in two tap halves so that only 23 float2
// of it are live next to the 31 float2 taps.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// HIN / HOUT: x / y are __half [tokens, 128] (fp32 arithmetic).  HOUT costs one F2FP per output pair and halves the write; HIN halves the
// staged tile (23.5 KB) but converts half2 -> float2 on every window load -- measured slower than the fp32 input (13.0 vs 10.0 ms per step at
// configs[1]: the kernel is bound by FFMA issue, not by its reads), so the model keeps u in fp32 and stores only v in fp16.
template <bool SWISH, bool HIN, bool HOUT = HIN>
__global__ void __launch_bounds__(128, 3) dwconv_bn_swish_kernel(const void* __restrict__ xv, const SebSeq sq,
                                                                const float* __restrict__ w, const float* __restrict__ bn_scale,
                                                                const float* __restrict__ bn_shift, void* __restrict__ yv, int nchunks) {
  using elem_t = typename std::conditional<HIN, __half, float>::type;
  using out_t = typename std::conditional<HOUT, __half, float>::type;
  const elem_t* __restrict__ x = reinterpret_cast<const elem_t*>(xv);
  out_t* __restrict__ y = reinterpret_cast<out_t*>(yv);
  __shared__ __align__(16) elem_t tile[DW_TI + DW_K - 1][DW_C];
  // position chunks are the fastest grid index: neighbouring chunks of a sequence run together, so the 30 halo rows they
  // share are served by L2 instead of a second trip to HBM
  const int seq = blockIdx.x / nchunks, i0 = (blockIdx.x - seq * nchunks) * DW_TI;
  const int cp = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const long long base = (long long)(seq / sq.inner) * sq.outer_stride + (seq % sq.inner);
  // stage rows i0-15 .. i0+64+15 (zero outside the sequence) with cp.async: 32 lanes x 16 B cover one 512-byte row and the
  // whole 47 KB tile is in flight at once (register-staged loads serialised four DRAM round trips per CTA)
  {
    constexpr int CPR = DW_C * (int)sizeof(elem_t) / 16;                  // 16-byte chunks per row: 32 (fp32) / 16 (fp16)
    constexpr int EPC = 16 / (int)sizeof(elem_t);                        // elements per chunk
    constexpr int NV = (DW_TI + DW_K - 1) * CPR;
    const uint32_t t0 = (uint32_t)__cvta_generic_to_shared(&tile[0][0]);
#pragma unroll 4
    for (int idx = threadIdx.x; idx < NV; idx += 128) {
      const int r = idx / CPR, c4 = idx % CPR;
      const int i = i0 + r - DW_PAD;
      const bool ok = i >= 0 && i < sq.n;
      const elem_t* src = x + (base + (long long)(ok ? i : 0) * sq.pos_stride) * DW_C + c4 * EPC;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(t0 + idx * 16), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float2 wr[DW_K];
#pragma unroll
  for (int k = 0; k < DW_K; ++k) wr[k] = __ldg(reinterpret_cast<const float2*>(w + k * DW_C) + cp);
  const float2 sc = __ldg(reinterpret_cast<const float2*>(bn_scale) + cp), sh = __ldg(reinterpret_cast<const float2*>(bn_shift) + cp);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  auto ld2 = [&](int row) -> float2 {       // channel pair cp of tile row `row`
    if (HIN) return __half22float2(*(reinterpret_cast<const __half2*>(&tile[0][0]) + row * (DW_C / 2) + cp));
    return *(reinterpret_cast<const float2*>(&tile[0][0]) + row * (DW_C / 2) + cp);
  };
#pragma unroll 1
  for (int gi = 0; gi < 4; ++gi) {
    const int gpos = ph * 32 + gi * 8;              // first output position of this group inside the tile
    if (i0 + gpos >= sq.n) break;
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = make_float2(0.f, 0.f);
    {   // taps 0..15 use window rows gpos .. gpos+22
      float2 win[23];
#pragma unroll
      for (int r = 0; r < 23; ++r) win[r] = ld2(gpos + r);
#pragma unroll
      for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = ffma2(wr[k], win[o + k], acc[o]);
    }
    {   // taps 16..30 use window rows gpos+16 .. gpos+37
      float2 win[22];
#pragma unroll
      for (int r = 0; r < 22; ++r) win[r] = ld2(gpos + 16 + r);
#pragma unroll
      for (int k = 16; k < DW_K; ++k)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = ffma2(wr[k], win[o + k - 16], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int i = i0 + gpos + o;
      if (i < sq.n) {
        float2 v = ffma2(acc[o], sc, sh);
        if (SWISH) { v.x *= sigmoidf_acc(v.x); v.y *= sigmoidf_acc(v.y); }
        if (HOUT) *reinterpret_cast<__half2*>(y + (base + (long long)i * sq.pos_stride) * DW_C + cp * 2) = __floats2half2_rn(v.x, v.y);
        else *reinterpret_cast<float2*>(y + (base + (long long)i * sq.pos_stride) * DW_C + cp * 2) = v;
      }
    }
  }
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_dwconv_bn_swish(const float* x, const SebSeq* seq, const float* w, const float* bn_scale, const float* bn_shift,
                                      float* y, void* stream) {
  SEB_REQUIRE(x && seq && w && bn_scale && bn_shift && y, SEB_EINVAL, "dwconv: null argument");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0, SEB_EINVAL, "dwconv: bad sequence descriptor");
  const int nchunks = (seq->n + DW_TI - 1) / DW_TI;
  const long long nblocks = (long long)seq->nseq * nchunks;
  SEB_REQUIRE(nblocks < 2147483647LL, SEB_EINVAL, "dwconv: grid too large");
  dwconv_bn_swish_kernel<true, false><<<(unsigned)nblocks, 128, 0, (cudaStream_t)stream>>>(x, *seq, w, bn_scale, bn_shift, y, nchunks);
  SEB_CHECK_LAUNCH("dwconv_bn_swish_kernel");
  return 0;
}

extern "C" int seb200_dwconv_bn_swish_f16(const void* x, int x_is_half, const SebSeq* seq, const float* w, const float* bn_scale, const float* bn_shift,
                                          void* y, void* stream) {
  SEB_REQUIRE(x && seq && w && bn_scale && bn_shift && y && aligned16(x) && aligned16(y), SEB_EINVAL, "dwconv (fp16): null / unaligned argument");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0, SEB_EINVAL, "dwconv: bad sequence descriptor");
  const int nchunks = (seq->n + DW_TI - 1) / DW_TI;
  const long long nblocks = (long long)seq->nseq * nchunks;
  SEB_REQUIRE(nblocks < 2147483647LL, SEB_EINVAL, "dwconv: grid too large");
  if (x_is_half) dwconv_bn_swish_kernel<true, true, true><<<(unsigned)nblocks, 128, 0, (cudaStream_t)stream>>>(x, *seq, w, bn_scale, bn_shift, y, nchunks);
  else dwconv_bn_swish_kernel<true, false, true><<<(unsigned)nblocks, 128, 0, (cudaStream_t)stream>>>(x, *seq, w, bn_scale, bn_shift, y, nchunks);
  SEB_CHECK_LAUNCH("dwconv_bn_swish_kernel<fp16 out>");
  return 0;
}

// The depthwise convolution alone, y = scale * DWConv31(x; w) + shift per channel (training step): with scale = 1,
// shift = conv bias it is the train-mode forward (BatchNorm1d then needs the batch statistics of y, conformer.py:166-167); with the taps
// reversed, scale = 1, shift = 0 it is the convolution's data gradient.  w [31][128] tap-major.
extern "C" int seb200_dwconv(const float* x, const SebSeq* seq, const float* w, const float* scale, const float* shift, float* y, void* stream) {
  SEB_REQUIRE(x && seq && w && scale && shift && y, SEB_EINVAL, "dwconv: null argument");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0, SEB_EINVAL, "dwconv: bad sequence descriptor");
  const int nchunks = (seq->n + DW_TI - 1) / DW_TI;
  const long long nblocks = (long long)seq->nseq * nchunks;
  SEB_REQUIRE(nblocks < 2147483647LL, SEB_EINVAL, "dwconv: grid too large");
  dwconv_bn_swish_kernel<false, false><<<(unsigned)nblocks, 128, 0, (cudaStream_t)stream>>>(x, *seq, w, scale, shift, y, nchunks);
  SEB_CHECK_LAUNCH("dwconv_kernel");
  return 0;
}
