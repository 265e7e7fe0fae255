// dwconv.cu -- DepthWiseConv1d(128, k=31, zero pad 15/15) + BatchNorm1d(eval, folded) + Swish along the sequence
// axis (conformer.py:166-168).  One CTA = 64 positions x 128 channels of one sequence; thread = channel; the
// (64+30) x 128 input tile is staged in shared memory (coalesced 512-byte rows), each thread produces 8 outputs per
// pass from a 38-value register window (38 LDS per 248 FMA).
#include "common.cuh"

namespace seb {

constexpr int DW_TI = 64, DW_K = 31, DW_PAD = 15, DW_C = 128;

__global__ void __launch_bounds__(128) dwconv_bn_swish_kernel(const float* __restrict__ x, const SebSeq sq,
                                                             const float* __restrict__ w, const float* __restrict__ bn_scale,
                                                             const float* __restrict__ bn_shift, float* __restrict__ y) {
  __shared__ float tile[DW_TI + DW_K - 1][DW_C];
  const int seq = blockIdx.x, i0 = blockIdx.y * DW_TI, c = threadIdx.x;
  const long long base = (long long)(seq / sq.inner) * sq.outer_stride + (seq % sq.inner);
  // stage rows i0-15 .. i0+64+15 (zero outside the sequence): 32 lanes x float4 cover one 512-byte row
  {
    constexpr int NV = (DW_TI + DW_K - 1) * (DW_C / 4);
    float4* t4 = reinterpret_cast<float4*>(&tile[0][0]);
#pragma unroll 6
    for (int idx = threadIdx.x; idx < NV; idx += 128) {
      const int r = idx >> 5, c4 = idx & 31;
      const int i = i0 + r - DW_PAD;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i >= 0 && i < sq.n) v = ldg4(x + (base + (long long)i * sq.pos_stride) * DW_C + c4 * 4);
      t4[idx] = v;
    }
  }
  float wr[DW_K];
#pragma unroll
  for (int k = 0; k < DW_K; ++k) wr[k] = __ldg(w + k * DW_C + c);
  const float sc = bn_scale[c], sh = bn_shift[c];
  __syncthreads();
#pragma unroll 1
  for (int g = 0; g < DW_TI; g += 8) {
    if (i0 + g >= sq.n) break;
    float win[8 + DW_K - 1];
#pragma unroll
    for (int r = 0; r < 8 + DW_K - 1; ++r) win[r] = tile[g + r][c];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < DW_K; ++k) acc = fmaf(wr[k], win[o + k], acc);
      const int i = i0 + g + o;
      if (i < sq.n) {
        float v = fmaf(acc, sc, sh);
        y[(base + (long long)i * sq.pos_stride) * DW_C + c] = v * sigmoidf_acc(v);
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
  dim3 grid(seq->nseq, (seq->n + DW_TI - 1) / DW_TI);   // sequences on x (2^31 limit), position chunks on y
  SEB_REQUIRE(grid.y <= 65535u, SEB_EINVAL, "dwconv: sequence too long");
  dwconv_bn_swish_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, *seq, w, bn_scale, bn_shift, y);
  SEB_CHECK_LAUNCH("dwconv_bn_swish_kernel");
  return 0;
}
