// mma_tf32.cuh -- warp-level fp32-grade matrix products on mma.sync m16n8k8 TF32 ("3xTF32": every fp32 operand is split into
// hi = tf32(x), lo = tf32(x - hi) and the product is accumulated as lo*hi + hi*lo + hi*hi in fp32, ~2^-20 relative operand error).
// Used by the training step's attention kernels and weight-gradient contractions, where gradients of the
// random-weight network amplify perturbations ~100x and a 10-bit operand mantissa is not enough.  Operands are fetched from shared
// memory through element accessors, so any layout / transposition is a one-line lambda at the call site.
//
// Fragment layouts (PTX ISA, m16n8k8 .tf32): g = lane >> 2, t = lane & 3
//   A (16 x 8, row):  a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)
//   B (8 x 8, col):   b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   C (16 x 8):       c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
#pragma once
#include "common.cuh"

namespace seb {
namespace tf32 {

// hi = x rounded to the 10-bit TF32 mantissa, lo = x - hi (exact in fp32) rounded the same way: |x - hi - lo| <= 2^-22 |x|, unbiased.
// The rounding is the integer form (add half an ulp of the 13 dropped bits, mask; a carry into the exponent is the correct result):
// cvt.rna.tf32.f32 is a multi-instruction sequence on sm_100 (FSETP / SEL / LOP3 / IMAD were 60 % of the first kernels' instruction
// stream, profiles/r2/ncu_full_attention_train_fwd.txt), and plain truncation of lo biases every product by 2^-21, which this network's
// ~100x gradient amplification turns into a measurable error (median gradient rel-L2 1.2e-3 instead of < 7.5e-4 at 4 x 2 s).
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[mt][nt] += A[16 mt .., k] * B[k, 8 nt ..] over k = 0 .. 8 KS - 1;  a_at(row, k) / b_at(k, col) return fp32 elements
template <int MT, int NT, class FA, class FB>
__device__ __forceinline__ void warp_gemm(float (&acc)[MT][NT][4], int KS, FA a_at, FB b_at) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = ks * 8;
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      split(a_at(mt * 16 + g, k0 + t), ah[mt][0], al[mt][0]);
      split(a_at(mt * 16 + g + 8, k0 + t), ah[mt][1], al[mt][1]);
      split(a_at(mt * 16 + g, k0 + t + 4), ah[mt][2], al[mt][2]);
      split(a_at(mt * 16 + g + 8, k0 + t + 4), ah[mt][3], al[mt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bh[2], bl[2];
      split(b_at(k0 + t, nt * 8 + g), bh[0], bl[0]);
      split(b_at(k0 + t + 4, nt * 8 + g), bh[1], bl[1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma(acc[mt][nt], al[mt], bh);
        mma(acc[mt][nt], ah[mt], bl);
        mma(acc[mt][nt], ah[mt], bh);
      }
    }
  }
}

// the same product restricted to the k-steps [ks_lo, ks_hi) and the n-tiles [nt_lo, nt_hi): fringe tiles (a sequence length of 64 k + 1 leaves ONE live
// row / column in the last tile) skip the dead part of the tensor work; the bounds are warp-uniform, the accumulator indexing stays static
template <int MT, int NT, class FA, class FB>
__device__ __forceinline__ void warp_gemm_range(float (&acc)[MT][NT][4], int ks_lo, int ks_hi, int nt_lo, int nt_hi, FA a_at, FB b_at) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int ks = ks_lo; ks < ks_hi; ++ks) {
    const int k0 = ks * 8;
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      split(a_at(mt * 16 + g, k0 + t), ah[mt][0], al[mt][0]);
      split(a_at(mt * 16 + g + 8, k0 + t), ah[mt][1], al[mt][1]);
      split(a_at(mt * 16 + g, k0 + t + 4), ah[mt][2], al[mt][2]);
      split(a_at(mt * 16 + g + 8, k0 + t + 4), ah[mt][3], al[mt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt >= nt_lo && nt < nt_hi) {
        uint32_t bh[2], bl[2];
        split(b_at(k0 + t, nt * 8 + g), bh[0], bl[0]);
        split(b_at(k0 + t + 4, nt * 8 + g), bh[1], bl[1]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma(acc[mt][nt], al[mt], bh);
          mma(acc[mt][nt], ah[mt], bl);
          mma(acc[mt][nt], ah[mt], bh);
        }
      }
    }
  }
}
// FR = false: the full product (KS k-steps, every n-tile), exactly warp_gemm; FR = true: the restricted one
template <bool FR, int MT, int NT, class FA, class FB>
__device__ __forceinline__ void warp_gemm_fr(float (&acc)[MT][NT][4], int KS, int ks_lo, int ks_hi, int nt_lo, int nt_hi, FA a_at, FB b_at) {
  if (FR) warp_gemm_range<MT, NT>(acc, ks_lo, ks_hi, nt_lo, nt_hi, a_at, b_at);
  else warp_gemm<MT, NT>(acc, KS, a_at, b_at);
}

template <int MT, int NT>
__device__ __forceinline__ void zero(float (&acc)[MT][NT][4]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
}

// row / column of accumulator element e of tile (mt, nt) for this lane
__device__ __forceinline__ int c_row(int mt, int e) { return mt * 16 + ((threadIdx.x & 31) >> 2) + ((e >> 1) << 3); }
__device__ __forceinline__ int c_col(int nt, int e) { return nt * 8 + ((threadIdx.x & 3) << 1) + (e & 1); }

}  // namespace tf32
}  // namespace seb
