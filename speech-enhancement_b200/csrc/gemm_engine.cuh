// gemm_engine.cuh -- the dense-contraction engine of libseb200.
//
//   C[M, N] = epilogue( loader(A)[M, K] * W[N, K]^T )
//
// Two main loops share the same operand loaders and epilogues:
//   * gemm_tc_kernel   : tcgen05.mma (kind::f16, bf16 operands, fp32 TMEM accumulator).  fp32
//                        activations are split on the fly into bf16 hi/lo and written to shared
//                        memory in the UMMA K-major SWIZZLE_128B canonical layout by producer
//                        warps; weights arrive pre-split / pre-swizzled by one cp.async.bulk
//                        per stage.  3 MMAs per K-step (hi*hi, hi*lo, lo*hi) give ~2^-17
//                        relative operand error ("BF16 3-product split").
//   * gemm_simt_kernel : plain fp32 FFMA tiles.  Bit-for-bit independent of the tensor path;
//                        used by tests to cross-check it and for the DFT where fp32 is needed.
#pragma once
#include "common.cuh"

namespace seb {

struct GemmArgs {
  const float* a[4];
  long long lda;
  const float* ln_g;
  const float* ln_b;
  int M, N, K;
  int B, T, Fin, Fout, taps_t, dil, stride_f, nslots;
  const float* bias;
  float* out;
  long long ldo;
  const float* resid;
  long long ldr;
  float alpha;
};

constexpr int BM = 128;   // rows (pixels / tokens / frames) per CTA tile
constexpr int BK = 64;    // K chunk: 64 bf16 = one 128-byte swizzle row

// -------------------------------------------------------------------------------------------
// Loaders.  256 producer threads cover a 128 x 64 fp32 tile in 4 passes: thread (rloc = tid>>3,
// sub = tid&7) loads 8 consecutive K values (32 bytes) of row pass*32 + rloc, so 8 adjacent
// lanes read one 256-byte row segment (coalesced) and own a full row for LayerNorm.
// -------------------------------------------------------------------------------------------
template <int KIND> struct Loader;

template <> struct Loader<SEB_LOAD_ROWS> {
  struct Row { const float* p; };
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) {
    r.p = (m < g.M) ? g.a[0] + (long long)m * g.lda : nullptr;
  }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    if (r.p) {
      const float* p = r.p + kc * BK + sub * 8;
      float4 x = ldg4(p), y = ldg4(p + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

// plain rows stored as __half (g.a[0] points at halfs, g.lda in halfs): exact in the bf16 hi | lo split (11-bit mantissa)
template <> struct Loader<SEB_LOAD_ROWS_F16> {
  struct Row { const __half* p; };
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) {
    r.p = (m < g.M) ? reinterpret_cast<const __half*>(g.a[0]) + (long long)m * g.lda : nullptr;
  }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    if (r.p) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(r.p + kc * BK + sub * 8));
      const uint32_t w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

// Two 64-wide row sources side by side (K = 128): K chunk 0 reads a[0], chunk 1 reads a[1] (MergeBlock: [x | conditioner],
// models/tsc_diffusion.py:32-34 -- merge_diffusion and conditioner_projection are one contraction over the pair)
template <> struct Loader<SEB_LOAD_ROWS2> {
  struct Row { long long off; };   // off < 0: row beyond M
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) { r.off = (m < g.M) ? (long long)m * g.lda : -1; }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    if (r.off >= 0) {
      const float* p = g.a[kc & 1] + r.off + sub * 8;
      float4 x = ldg4(p), y = ldg4(p + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

// LayerNorm(64, eps 1e-5) fused into the load (PreNorm, conformer.py:63-71 and net[0] of the conv module)
template <> struct Loader<SEB_LOAD_ROWS_LN> {
  using Row = Loader<SEB_LOAD_ROWS>::Row;
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) { Loader<SEB_LOAD_ROWS>::init_row(g, m, r); }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    Loader<SEB_LOAD_ROWS>::load(g, r, kc, sub, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mean = s * (1.0f / 64.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 4);
    const float rstd = 1.0f / sqrtf(q * (1.0f / 64.0f) + 1e-5f);
    float4 g0 = ldg4(g.ln_g + sub * 8), g1 = ldg4(g.ln_g + sub * 8 + 4);
    float4 b0 = ldg4(g.ln_b + sub * 8), b1 = ldg4(g.ln_b + sub * 8 + 4);
    v[0] = v[0] * rstd * g0.x + b0.x; v[1] = v[1] * rstd * g0.y + b0.y;
    v[2] = v[2] * rstd * g0.z + b0.z; v[3] = v[3] * rstd * g0.w + b0.w;
    v[4] = v[4] * rstd * g1.x + b1.x; v[5] = v[5] * rstd * g1.y + b1.y;
    v[6] = v[6] * rstd * g1.z + b1.z; v[7] = v[7] * rstd * g1.w + b1.w;
  }
};

// Implicit-GEMM Conv2d on channels-last slots.  K order = (kt, kf, slot, 64 channels); the input
// of dense layer i is the list of 64-channel tensors [out_{i-1}, ..., out_1, x] (generator.py:31)
// addressed through g.a[slot] -- the concat is never materialised.
template <> struct Loader<SEB_LOAD_CONV> {
  struct Row { int b, t, f; };   // b < 0: row beyond M
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) {
    if (m < g.M) {
      int bt = m / g.Fout;
      r.f = m - bt * g.Fout;
      r.b = bt / g.T;
      r.t = bt - r.b * g.T;
    } else { r.b = -1; r.t = 0; r.f = 0; }
  }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    const int tap = kc / g.nslots;
    const int slot = kc - tap * g.nslots;
    const int kt = (g.taps_t == 2) ? tap / 3 : 0;
    const int kf = tap - kt * 3;
    const int tt = r.t - (g.taps_t - 1 - kt) * g.dil;
    const int ff = r.f * g.stride_f + kf - 1;
    if (r.b >= 0 && tt >= 0 && ff >= 0 && ff < g.Fin) {
      const float* p = g.a[slot] + (((long long)r.b * g.T + tt) * g.Fin + ff) * 64 + sub * 8;
      float4 x = ldg4(p), y = ldg4(p + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

// Adjoint of the implicit-GEMM convolution (training: dgrad of DilatedDenseNet / conv_2 / SPConvTranspose2d).
// Rows are the pixels (b, t', f') of the forward conv's INPUT (M = B * T * g.Fout, g.Fout = forward input width); the A operand is the
// gradient image of the forward conv's OUTPUT: [B, T, g.Fin, g.lda channels] (g.Fin = forward output width, g.lda = 64 or 128).
// K order = (tap = kt * 3 + kf, 64-channel sub-chunk): chunk kc reads dY[b, t' + (taps_t - 1 - kt) * dil, (f' + 1 - kf) / stride_f, sub * 64 ..]
// when that pixel exists (the forward conv read X[t - (taps_t - 1 - kt) * dil, fo * stride_f + kf - 1], so X[t', f'] fed exactly those outputs).
template <> struct Loader<SEB_LOAD_CONV_ADJ> {
  using Row = Loader<SEB_LOAD_CONV>::Row;
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) { Loader<SEB_LOAD_CONV>::init_row(g, m, r); }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    const int tap = kc / g.nslots;
    const int part = kc - tap * g.nslots;
    const int kt = (g.taps_t == 2) ? tap / 3 : 0;
    const int kf = tap - kt * 3;
    const int tt = r.t + (g.taps_t - 1 - kt) * g.dil;
    const int num = r.f + 1 - kf;
    const int ff = num / g.stride_f;
    if (r.b >= 0 && tt < g.T && num >= 0 && ff * g.stride_f == num && ff < g.Fin) {
      const float* p = g.a[0] + (((long long)r.b * g.T + tt) * g.Fin + ff) * g.lda + part * 64 + sub * 8;
      float4 x = ldg4(p), y = ldg4(p + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

// STFT framing: row (b, t) is the overlapping window xpad[b, t*hop : t*hop + n_fft] (no copy).
// g.Fin = n_fft (400), g.stride_f = hop (100), g.T = frames per utterance, g.lda = samples per row of xpad.
template <> struct Loader<SEB_LOAD_HANKEL> {
  struct Row { const float* p; };
  __device__ static void init_row(const GemmArgs& g, int m, Row& r) {
    if (m < g.M) {
      int b = m / g.T, t = m - b * g.T;
      r.p = g.a[0] + (long long)b * g.lda + (long long)t * g.stride_f;
    } else r.p = nullptr;
  }
  __device__ static void load(const GemmArgs& g, const Row& r, int kc, int sub, float (&v)[8]) {
    const int k = kc * BK + sub * 8;
    if (r.p && k < g.Fin) {
      float4 x = ldg4(r.p + k), y = ldg4(r.p + k + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

// -------------------------------------------------------------------------------------------
// Epilogues: called with 4 consecutive output columns n..n+3 (n % 4 == 0) of row m.
// -------------------------------------------------------------------------------------------
template <int KIND> struct Epi;

__device__ __forceinline__ float4 add_bias(const GemmArgs& g, int n, float4 v) {
  if (g.bias) { float4 b = ldg4(g.bias + n); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
  return v;
}

template <> struct Epi<SEB_EPI_BIAS> {
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    if (n + 4 <= g.N) {
      st4(g.out + (long long)m * g.ldo + n, add_bias(g, n, v));
    } else {   // ragged tail (N % 4 != 0 never carries a bias on this path)
      float t[4] = {v.x, v.y, v.z, v.w};
      for (int i = 0; i < 4 && n + i < g.N; ++i) g.out[(long long)m * g.ldo + n + i] = t[i] + (g.bias ? g.bias[n + i] : 0.f);
    }
  }
};

template <> struct Epi<SEB_EPI_SWISH> {
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    v.x *= sigmoidf_acc(v.x); v.y *= sigmoidf_acc(v.y); v.z *= sigmoidf_acc(v.z); v.w *= sigmoidf_acc(v.w);
    st4(g.out + (long long)m * g.ldo + n, v);
  }
};

template <> struct Epi<SEB_EPI_GLU> {   // packed columns: (value_j, gate_j) adjacent (conformer.py:36-37)
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    float2 o = make_float2(v.x * sigmoidf_acc(v.y), v.z * sigmoidf_acc(v.w));
    *reinterpret_cast<float2*>(g.out + (long long)m * g.ldo + (n >> 1)) = o;
  }
};

template <> struct Epi<SEB_EPI_GLU_F16> {   // GLU with a __half output [M, N / 2]
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    const __half2 o = __floats2half2_rn(v.x * sigmoidf_acc(v.y), v.z * sigmoidf_acc(v.w));
    *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(g.out) + (long long)m * g.ldo + (n >> 1)) = o;
  }
};

template <> struct Epi<SEB_EPI_RESID> {
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    float4 r = *reinterpret_cast<const float4*>(g.resid + (long long)m * g.ldr + n);
    v.x = g.alpha * v.x + r.x; v.y = g.alpha * v.y + r.y; v.z = g.alpha * v.z + r.z; v.w = g.alpha * v.w + r.w;
    st4(g.out + (long long)m * g.ldo + n, v);
  }
};

// MergeBlock gate (models/tsc_diffusion.py:36-37): packed columns (gate_j, filter_j); the diffusion-step projection enters as one
// extra bias row per group of g.ldr consecutive rows (g.resid [groups, N]; null: none)
template <> struct Epi<SEB_EPI_GATE> {
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    if (g.resid) {
      const float4 r = ldg4(g.resid + (long long)(m / (int)g.ldr) * g.N + n);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    float2 o = make_float2(sigmoidf_acc(v.x) * tanhf_acc(v.y), sigmoidf_acc(v.z) * tanhf_acc(v.w));
    *reinterpret_cast<float2*>(g.out + (long long)m * g.ldo + (n >> 1)) = o;
  }
};

template <> struct Epi<SEB_EPI_RESID_SCALE> {   // (x + output_residual(y)) / sqrt(2): tsc_diffusion.py:39-41
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    float4 r = *reinterpret_cast<const float4*>(g.resid + (long long)m * g.ldr + n);
    v.x = g.alpha * (v.x + r.x); v.y = g.alpha * (v.y + r.y); v.z = g.alpha * (v.z + r.z); v.w = g.alpha * (v.w + r.w);
    st4(g.out + (long long)m * g.ldo + n, v);
  }
};

template <> struct Epi<SEB_EPI_SUBPIXEL> {   // generator.py:88-91: out[b, c, t, 2w + r] = y[b, r*64 + c, t, w]
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    v = add_bias(g, n, v);
    const int bt = m / g.Fout, w = m - bt * g.Fout;
    const int r = n >> 6, c = n & 63;
    st4(g.out + (((long long)bt * (2 * g.Fout)) + 2 * w + r) * 64 + c, v);
  }
};

template <> struct Epi<SEB_EPI_COMPRESS> {   // core/function.py:625-634 fused: columns are (re_k, im_k) pairs
  __device__ static void one(float* o, float re, float im) {
    const float m2 = re * re + im * im;
    float magc = 0.f, s = 0.f;
    if (m2 > 0.f) {
      const float r = sqrtf(m2);
      magc = powf(r, 0.3f);
      s = magc / r;
    }
    o[0] = magc; o[1] = re * s; o[2] = im * s;
  }
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    float* o = g.out + (long long)m * g.ldo + (n >> 1) * 3;
    one(o, v.x, v.y);
    if (n + 2 < g.N) one(o + 3, v.z, v.w);
  }
};

// q | k | v projection for the tensor-core attention: fp16 output [tokens, 192]; the q third is pre-scaled by
// dim_head^-0.5 * log2(e) (conformer.py:103,110 scale both logit terms by dim_head^-0.5; softmax is evaluated base 2)
template <> struct Epi<SEB_EPI_QKV_F16> {
  __device__ static void apply(const GemmArgs& g, int m, int n, float4 v) {
    if (m >= g.M || n >= g.N) return;
    const float sc = (n < 64) ? 0.25f * 1.4426950408889634f : 1.0f;
    __half2 a = __floats2half2_rn(v.x * sc, v.y * sc), b = __floats2half2_rn(v.z * sc, v.w * sc);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(g.out) + (long long)m * g.ldo + n) = pk;
  }
};

// -------------------------------------------------------------------------------------------
// SIMT main loop (fp32 FFMA).  grid = (ceil(M/128), npad/64), 256 threads, 8x4 outputs / thread.
// -------------------------------------------------------------------------------------------
constexpr int SIMT_LDA = BM + 4;
constexpr int SIMT_SMEM = (BK * SIMT_LDA + BK * 64) * (int)sizeof(float);

template <int LK, int EK>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g, const float* __restrict__ w, int npad) {
  extern __shared__ __align__(16) float smem_f[];
  float (*As)[SIMT_LDA] = reinterpret_cast<float (*)[SIMT_LDA]>(smem_f);          // [k][m]
  float (*Ws)[64] = reinterpret_cast<float (*)[64]>(smem_f + BK * SIMT_LDA);      // [k][n]
  const int tid = threadIdx.x, sub = tid & 7, rloc = tid >> 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * 64;
  typename Loader<LK>::Row rows[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) Loader<LK>::init_row(g, m0 + p * 32 + rloc, rows[p]);
  const int tm = tid >> 4, tn = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nkc = g.K / BK;
  for (int kc = 0; kc < nkc; ++kc) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[8];
      Loader<LK>::load(g, rows[p], kc, sub, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) As[sub * 8 + i][p * 32 + rloc] = v[i];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + j * 256, kr = idx >> 4, nc = (idx & 15) * 4;
      *reinterpret_cast<float4*>(&Ws[kr][nc]) = ldg4(w + (long long)(kc * BK + kr) * npad + n0 + nc);
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    Epi<EK>::apply(g, m0 + tm * 8 + i, n0 + tn * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
}

// -------------------------------------------------------------------------------------------
// tcgen05 main loop
// -------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch error, never as a hung GPU.
#ifndef SEB_MBAR_TESTWAIT
#define SEB_MBAR_TESTWAIT 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if SEB_MBAR_TESTWAIT
  // non-blocking poll: test_wait returns immediately, so the waiter reacts within a few cycles of the arrival
  for (uint32_t it = 0; it < (1u << 30); ++it)
    if (mbar_test_wait(bar, parity)) return;
  __trap();
#else
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
#endif
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major, SWIZZLE_128B canonical operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for K-major swizzle)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
}  // namespace ptx

constexpr int TC_THREADS = 320;        // default shape: 8 producer/epilogue warps + MMA warp + weight-copy warp
constexpr int TC_A_BYTES = BM * 128;   // one bf16 plane of the A tile

template <int NT, int NPL = 2> constexpr int tc_stage_bytes() { return NPL * TC_A_BYTES + NPL * NT * 128; }
template <int NT, int STAGES, int NPL = 2> constexpr int tc_smem_bytes() { return STAGES * tc_stage_bytes<NT, NPL>() + 1024; }
template <int NT> constexpr uint32_t tc_tmem_cols() { return NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256; }

// PW = producer/epilogue warps (8 or 4), MINB = CTAs per SM the register budget is sized for.  The K = 64 / 128 token
// GEMMs are bandwidth/latency bound: they run as PW = 4, one stage, n-tile 64 -> 49 KB smem, 64 TMEM columns and
// 192 threads per CTA, i.e. 4 resident CTAs per SM keep four tiles' loads in flight.
// NPL = operand planes.  2: x = hi + lo, products hi*hi + hi*lo + lo*hi (network GEMMs, ~2^-17 operand error).
// 3: x = hi + mid + lo, six products (all terms down to 2^-24): fp32-grade accuracy on the tensor pipe, used for the
// DFT / iDFT, where |X|^0.3 amplifies operand rounding on near-zero bins.
template <int NT, int STAGES, int LK, int EK, int PW = 8, int MINB = 2, int NPL = 2>
__global__ void __launch_bounds__((PW + 2) * 32, MINB)
gemm_tc_kernel(const GemmArgs g, const uint8_t* __restrict__ w_tc) {
  static_assert(PW == 4 || PW == 8, "producer warps must cover the four TMEM lane quarters once or twice");
  static_assert(NPL == 2 || NPL == 3, "two or three bf16 planes per operand");
  static_assert(NT % 16 == 0 && NT >= 16 && NT <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS/LDS, not generic ST/LD)
  const uint32_t smem_base = ptx::smem_u32(smem);
  constexpr int STAGE = tc_stage_bytes<NT, NPL>();
  constexpr uint32_t W_BYTES = NT * 128;
  // NPL == 3 keeps TWO accumulators: the hi x hi products go to the first, the five cross terms (2^-8 .. 2^-16 of it) to the second, and the
  // epilogue adds them.  The tensor pipe's fp32 accumulation truncates, so every MMA that lands in an accumulator costs ~half an ulp OF THAT
  // ACCUMULATOR, biased toward zero: with all six products in one accumulator a K = 1536 conv took 576 such steps (1.8e-5 of peak after the
  // whole generator); split this way the large accumulator sees K / 16 steps and the small one's ulp is 2^-8 of it.
  // NT == 64 with three planes (the training step's convolutions, data gradients and projections) folds the weight planes into N: they are consecutive
  // 64-row groups of ONE K-major operand, so a_hi x [w_hi | w_mid | w_lo] (N = 192), a_mid x [w_hi | w_mid] (N = 128) and a_lo x w_hi (N = 64) are three
  // instructions instead of six per k-step, every A plane is read from shared memory once instead of 3 / 2 / 1 times, and the column bases make the
  // equal-magnitude terms share THREE accumulators: hi.hi | hi.mid + mid.hi | hi.lo + mid.mid + lo.hi.
  constexpr bool FOLD = NPL == 3 && NT == 64;
  constexpr uint32_t ACC_COLS = tc_tmem_cols<NT>();
  constexpr uint32_t TMEM_COLS = FOLD ? 256 : (NPL == 3 ? 2 : 1) * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "two accumulators exceed tensor memory");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int ntile = blockIdx.y;
  const int nkc = g.K / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], PW * 32 + 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(&accum_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == PW) ptx::tmem_alloc(&tmem_base_s, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < PW) {
    // ---------------- producers: fp32 global -> bf16 hi/lo swizzled shared ----------------
    const int sub = tid & 7, rloc = tid >> 3;
    constexpr int RPP = PW * 4, NPASS = BM / RPP;      // rows per pass, passes per tile
    typename Loader<LK>::Row rows[NPASS];
#pragma unroll
    for (int p = 0; p < NPASS; ++p) Loader<LK>::init_row(g, m0 + p * RPP + rloc, rows[p]);
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc % STAGES;
      const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
      ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* a_hi = smem + s * STAGE;
      uint8_t* a_lo = a_hi + (NPL - 1) * TC_A_BYTES;      // planes: hi | (mid |) lo
#pragma unroll
      for (int p = 0; p < NPASS; ++p) {
        float v[8];
        Loader<LK>::load(g, rows[p], kc, sub, v);
        const int r = p * RPP + rloc;
        const int off = r * 128 + ((sub ^ (r & 7)) << 4);
        uint4 hi, lo;
        if (NPL == 2) {
          split_bf16x2(v[0], v[1], hi.x, lo.x);
          split_bf16x2(v[2], v[3], hi.y, lo.y);
          split_bf16x2(v[4], v[5], hi.z, lo.z);
          split_bf16x2(v[6], v[7], hi.w, lo.w);
        } else {
          uint4 mid;
          split3_bf16x2(v[0], v[1], hi.x, mid.x, lo.x);
          split3_bf16x2(v[2], v[3], hi.y, mid.y, lo.y);
          split3_bf16x2(v[4], v[5], hi.z, mid.z, lo.z);
          split3_bf16x2(v[6], v[7], hi.w, mid.w, lo.w);
          *reinterpret_cast<uint4*>(a_hi + TC_A_BYTES + off) = mid;
        }
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full_bar[s]);
    }
    // ---------------- epilogue: TMEM -> registers -> (warp-private smem transpose) -> global ----------------
    // tcgen05.ld 32x32b hands every lane one accumulator ROW; storing that way makes each warp-wide store touch 32
    // different rows (32 half-used sectors per instruction).  The rows are therefore bounced through a 4 KB
    // XOR-swizzled staging tile in the (now idle) stage-0 operand buffer so that 8 consecutive lanes write 128
    // contiguous bytes of one row and the epilogue functors see a coalesced (row, column) mapping.
    ptx::mbar_wait(&accum_bar, 0);
    ptx::tc_fence_after();
    // every producer's operand stores into the ring precede its last full_bar arrival, and accum_bar completes only after the MMAs that read
    // them: the ring is free for the staging tile.  The named barrier restates that order in a form compute-sanitizer's racecheck models
    // (it tracks bar.sync, not mbarrier chains); a few cycles per CTA.
    asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory");
    const int wq = warp & 3, half = warp >> 2;
    constexpr int HALF_COLS = NT / (PW / 4);     // columns per epilogue warp; multiple of 8
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * HALF_COLS);
    float4* stg = reinterpret_cast<float4*>(smem + warp * 4096);    // [32 rows][8 x float4]
#pragma unroll 1
    for (int c0 = 0; c0 < HALF_COLS; c0 += 32) {
      const int ncols = (HALF_COLS - c0 < 32) ? HALF_COLS - c0 : 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < ncols) {
          float v[8];
          ptx::tmem_ld8(taddr + c0 + j, v);
          if (FOLD) {
            float v2[8], v3[8];
            ptx::tmem_ld8(taddr + 64 + c0 + j, v2);
            ptx::tmem_ld8(taddr + 128 + c0 + j, v3);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += v3[i] + v2[i];
          } else if (NPL == 3) {
            float v2[8];
            ptx::tmem_ld8(taddr + ACC_COLS + c0 + j, v2);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += v2[i];
          }
          stg[lane * 8 + (((j >> 2) + 0) ^ (lane & 7))] = make_float4(v[0], v[1], v[2], v[3]);
          stg[lane * 8 + (((j >> 2) + 1) ^ (lane & 7))] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      __syncwarp();
      const int ch = lane & 7;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int R = it * 4 + (lane >> 3);
        if (ch * 4 < ncols) {
          const float4 val = stg[R * 8 + (ch ^ (R & 7))];
          Epi<EK>::apply(g, m0 + wq * 32 + R, ntile * NT + half * HALF_COLS + c0 + ch * 4, val);
        }
      }
      __syncwarp();
    }
    ptx::tc_fence_before();
  } else if (warp == PW) {
    // ---------------- MMA issuer (one thread) ----------------
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kc = 0; kc < nkc; ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t base = smem_base + s * STAGE;
        const uint64_t a_hi = ptx::umma_desc_sw128(base);
        const uint64_t a_lo = ptx::umma_desc_sw128(base + (NPL - 1) * TC_A_BYTES);
        const uint64_t w_hi = ptx::umma_desc_sw128(base + NPL * TC_A_BYTES);
        const uint64_t w_lo = ptx::umma_desc_sw128(base + NPL * TC_A_BYTES + (NPL - 1) * W_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ko = (uint64_t)((k * 32) >> 4);   // 16 bf16 = 32 bytes along K inside the swizzle row
          if (FOLD) {
            constexpr uint32_t IDB = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t first = (kc | k) ? 1u : 0u;
            const uint64_t a_mid = ptx::umma_desc_sw128(base + TC_A_BYTES);
            ptx::mma_bf16(tmem_base, a_hi + ko, w_hi + ko, IDB | ((uint32_t)(192 >> 3) << 17), first);
            ptx::mma_bf16(tmem_base + 64u, a_mid + ko, w_hi + ko, IDB | ((uint32_t)(128 >> 3) << 17), 1u);
            ptx::mma_bf16(tmem_base + 128u, a_lo + ko, w_hi + ko, IDB | ((uint32_t)(64 >> 3) << 17), 1u);
          } else if (NPL == 3) {
            const uint32_t first = (kc | k) ? 1u : 0u;
            const uint32_t acc2 = tmem_base + ACC_COLS;                                    // cross terms: their own accumulator (see above)
            const uint64_t a_mid = ptx::umma_desc_sw128(base + TC_A_BYTES);
            const uint64_t w_mid = ptx::umma_desc_sw128(base + NPL * TC_A_BYTES + W_BYTES);
            ptx::mma_bf16(acc2, a_lo + ko, w_hi + ko, IDESC, first);                       // smallest terms first
            ptx::mma_bf16(acc2, a_hi + ko, w_lo + ko, IDESC, 1u);
            ptx::mma_bf16(acc2, a_mid + ko, w_mid + ko, IDESC, 1u);
            ptx::mma_bf16(acc2, a_mid + ko, w_hi + ko, IDESC, 1u);
            ptx::mma_bf16(acc2, a_hi + ko, w_mid + ko, IDESC, 1u);
            ptx::mma_bf16(tmem_base, a_hi + ko, w_hi + ko, IDESC, first);
          } else {
            ptx::mma_bf16(tmem_base, a_lo + ko, w_hi + ko, IDESC, (kc | k) ? 1u : 0u);      // smallest terms first
            ptx::mma_bf16(tmem_base, a_hi + ko, w_lo + ko, IDESC, 1u);
            ptx::mma_bf16(tmem_base, a_hi + ko, w_hi + ko, IDESC, 1u);
          }
        }
        ptx::tc_commit(&empty_bar[s]);
      }
      ptx::tc_commit(&accum_bar);
    }
  } else {
    // ---------------- weight stager: one bulk copy (hi|lo image) per stage ----------------
    if (lane == 0) {
      const uint8_t* src = w_tc + (size_t)ntile * nkc * (NPL * W_BYTES);
      for (int kc = 0; kc < nkc; ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full_bar[s], NPL * W_BYTES);
        ptx::bulk_g2s(smem_base + s * STAGE + NPL * TC_A_BYTES, src + (size_t)kc * (NPL * W_BYTES), NPL * W_BYTES, &full_bar[s]);
      }
    }
  }
  __syncthreads();
  if (warp == PW) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// -------------------------------------------------------------------------------------------
// Implicit-GEMM convolution on PRE-SPLIT activations (SEB_LOAD_CONV_SPLIT).
//
// The producers of conv inputs (inorm_prelu / split_planes) store every pixel as 256 bytes: 64 bf16 `hi` then
// 64 bf16 `lo` with hi + lo == x to 2^-17 -- the same bytes as fp32, but already the two MMA operands.  The loader
// is then a pure copy: 16-byte cp.async (zero-filled outside the image) straight into the swizzled UMMA tile, no
// register pass, no conversion; a software-pipelined wait_group + fence.proxy.async hands each stage to the
// tensor pipe.  Same mbarrier ring, weight stager, MMA issuer and epilogue as gemm_tc_kernel.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

template <int NT, int STAGES, int EK>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_split_tc_kernel(const GemmArgs g, const uint8_t* __restrict__ w_tc) {
  static_assert(STAGES == 1 || STAGES == 2, "producer lag is STAGES - 1 with a literal wait_group");
  static_assert(NT % 16 == 0 && NT >= 16 && NT <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS/LDS, not generic ST/LD)
  const uint32_t smem_base = ptx::smem_u32(smem);
  constexpr int STAGE = tc_stage_bytes<NT>();
  constexpr uint32_t W_BYTES = NT * 128;
  constexpr uint32_t TMEM_COLS = tc_tmem_cols<NT>();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int ntile = blockIdx.y;
  const int nkc = g.K / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 8 / STAGES + 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(&accum_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc(&tmem_base_s, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ---------------- producers: cp.async of pre-split bf16 planes ----------------
    // The 8 producer warps form STAGES independent groups; group gidx owns ring slot gidx and loads every STAGES-th
    // K-chunk into it: wait for the slot, copy, wait for the copy, publish.  Publishing chunk c therefore depends only
    // on MMA(c - STAGES) and on its own copy -- never on another chunk's slot.  (With one group that published chunk
    // c - 1 after queueing chunk c, the publish sat behind the wait for MMA(c - 1): a fully serialised pipeline,
    // measured 1.2 k cycles per 384-cycle chunk even with the copies disabled.)
    // lane -> (16-byte chunk c = lane & 7, plane = (lane >> 3) & 1, row-in-pair = lane >> 4): one warp instruction
    // copies 2 pixels x 256 contiguous bytes.
    constexpr int GW = 8 / STAGES;                       // warps per group
    constexpr int RPP = GW * 2, NPASS = BM / RPP;        // rows per pass, passes per chunk
    const int gidx = warp / GW, gw = warp % GW;
    const int c = lane & 7, plane = (lane >> 3) & 1, r0 = gw * 2 + (lane >> 4);
    const uint32_t dst0 = (uint32_t)(plane * TC_A_BYTES + r0 * 128 + ((c ^ (r0 & 7)) << 4));   // + p * RPP * 128 per pass (RPP % 8 == 0)
    const int src_lane_off = plane * 128 + c * 16;                                             // bytes inside a pixel
    int pix[NPASS];           // flat input pixel index of the centre tap of this thread's row in pass p (-1: row beyond M)
    int tf[NPASS];            // (t << 16) | f
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      const int m = m0 + p * RPP + r0;
      if (m < g.M) {
        const int bt = m / g.Fout, f = m - bt * g.Fout, b = bt / g.T, t = bt - b * g.T;
        pix[p] = (b * g.T + t) * g.Fin + f * g.stride_f;
        tf[p] = (t << 16) | f;
      } else { pix[p] = -1; tf[p] = 0; }
    }
    for (int kc = gidx; kc < nkc; kc += STAGES) {
      const int s = gidx;
      const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
      const int tap = kc / g.nslots, slot = kc - tap * g.nslots;
      const int kt = (g.taps_t == 2) ? tap / 3 : 0, kf = tap - kt * 3;
      const int dt = (g.taps_t - 1 - kt) * g.dil, df = kf - 1;
      const uint8_t* src_base = reinterpret_cast<const uint8_t*>(g.a[slot]) + src_lane_off;
      const int dpix = df - dt * g.Fin;
      const uint32_t dst = smem_base + s * STAGE + dst0;
      ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
#pragma unroll
      for (int p = 0; p < NPASS; ++p) {
        const int t = tf[p] >> 16, f = tf[p] & 0xffff;
        const int ff = f * g.stride_f + df;
        const bool ok = pix[p] >= 0 && t >= dt && ff >= 0 && ff < g.Fin;
        const long long q = ok ? (long long)(pix[p] + dpix) : 0;
        cp_async16_zfill(dst + p * (RPP * 128), src_base + q * 256, ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_bar[s]);
    }

    // ---------------- epilogue (identical to gemm_tc_kernel) ----------------
    ptx::mbar_wait(&accum_bar, 0);
    ptx::tc_fence_after();
    asm volatile("bar.sync 1, 256;" ::: "memory");     // see gemm_tc_kernel: orders the ring's cp.async writes before the staging stores for racecheck
    const int wq = warp & 3, half = warp >> 2;
    constexpr int HALF_COLS = NT / 2;
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * HALF_COLS);
    float4* stg = reinterpret_cast<float4*>(smem + warp * 4096);
#pragma unroll 1
    for (int c0 = 0; c0 < HALF_COLS; c0 += 32) {
      const int ncols = (HALF_COLS - c0 < 32) ? HALF_COLS - c0 : 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < ncols) {
          float v[8];
          ptx::tmem_ld8(taddr + c0 + j, v);
          stg[lane * 8 + (((j >> 2) + 0) ^ (lane & 7))] = make_float4(v[0], v[1], v[2], v[3]);
          stg[lane * 8 + (((j >> 2) + 1) ^ (lane & 7))] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      __syncwarp();
      const int ch = lane & 7;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int R = it * 4 + (lane >> 3);
        if (ch * 4 < ncols) {
          const float4 val = stg[R * 8 + (ch ^ (R & 7))];
          Epi<EK>::apply(g, m0 + wq * 32 + R, ntile * NT + half * HALF_COLS + c0 + ch * 4, val);
        }
      }
      __syncwarp();
    }
    ptx::tc_fence_before();
  } else if (warp == 8) {
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kc = 0; kc < nkc; ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t base = smem_base + s * STAGE;
        const uint64_t a_hi = ptx::umma_desc_sw128(base);
        const uint64_t a_lo = ptx::umma_desc_sw128(base + TC_A_BYTES);
        const uint64_t w_hi = ptx::umma_desc_sw128(base + 2 * TC_A_BYTES);
        const uint64_t w_lo = ptx::umma_desc_sw128(base + 2 * TC_A_BYTES + W_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ko = (uint64_t)((k * 32) >> 4);
          ptx::mma_bf16(tmem_base, a_lo + ko, w_hi + ko, IDESC, (kc | k) ? 1u : 0u);
          ptx::mma_bf16(tmem_base, a_hi + ko, w_lo + ko, IDESC, 1u);
          ptx::mma_bf16(tmem_base, a_hi + ko, w_hi + ko, IDESC, 1u);
        }
        ptx::tc_commit(&empty_bar[s]);
      }
      ptx::tc_commit(&accum_bar);
    }
  } else {
    if (lane == 0) {
      const uint8_t* src = w_tc + (size_t)ntile * nkc * (2 * W_BYTES);
      for (int kc = 0; kc < nkc; ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (uint32_t)(kc / STAGES) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * W_BYTES);
        ptx::bulk_g2s(smem_base + s * STAGE + 2 * TC_A_BYTES, src + (size_t)kc * (2 * W_BYTES), 2 * W_BYTES, &full_bar[s]);
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace seb
