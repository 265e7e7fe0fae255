// train_elem.cu -- token- / pixel-wise kernels of the generator's TRAINING step: train-mode forward pieces the fused
// inference kernels cannot provide (dropout at /root/reference/models/conformer.py:125,139,141; BatchNorm1d batch statistics + running
// update at :167) and the backward of every non-GEMM op of the path (LayerNorm, Swish, GLU, BatchNorm, InstanceNorm2d + PReLU, the two
// decoder heads, the mask tail and the recombination; loss.backward() at core/function.py:274).  All HBM-bound, fp32, channels-last.
//
// Reductions over tokens / pixels (parameter gradients, normalisation sums) are two-stage and deterministic: every CTA writes one row of
// partial sums, a finish kernel adds the rows in a fixed order (no atomics).  Sums that feed a normalisation are carried in fp64 like the
// forward statistics; BatchNorm / its backward expose their LOCAL sums so that SyncBatchNorm semantics (main_gan.py:154-155) are one small
// all-reduce of that vector between the `sums` and the `apply` call.
#include "common.cuh"

namespace seb {

static int tgrid(long long items, int per_block, int cap = 148 * 8) {
  long long g = (items + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ float swish_grad(float a) {            // d/da [a * sigmoid(a)]
  const float s = sigmoidf_acc(a);
  return s * fmaf(a, 1.0f - s, 1.0f);
}

// ---- Philox4x32-10 keep-mask: mask[i] = (u_i >= p), u_i uniform in [0, 1) from counter (offset + i / 4), key = seed ----------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(unsigned char* __restrict__ mask, long long n4, long long n, float p,
                                                          unsigned long long seed, unsigned long long offset) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const unsigned long long c = offset + (unsigned long long)i;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    unsigned char k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = ((float)(rr[j] >> 8) * (1.0f / 16777216.0f) >= p) ? 1 : 0;
    if (i * 4 + 3 < n) {
      *reinterpret_cast<uchar4*>(mask + i * 4) = make_uchar4(k[0], k[1], k[2], k[3]);
    } else {
      for (int j = 0; j < 4 && i * 4 + j < n; ++j) mask[i * 4 + j] = k[j];
    }
  }
}

// ---- elementwise: 4 values per thread ---------------------------------------------------------------------------------------------
// MODE 0: h = swish(a) * keep * scale             1: da = dh * keep * scale * swish'(a)
// MODE 2: y = resid + scale * keep * t            3: dt = scale * keep * dy
template <int MODE>
__global__ void __launch_bounds__(256) elem_mask_kernel(const float* __restrict__ a, const float* __restrict__ b, const unsigned char* __restrict__ mask,
                                                       float scale, float* __restrict__ out, long long n4) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 x = ldg4(a + i * 4);
    float k[4] = {scale, scale, scale, scale};
    if (mask) {
      const uchar4 m = *reinterpret_cast<const uchar4*>(mask + i * 4);
      k[0] = m.x ? scale : 0.f; k[1] = m.y ? scale : 0.f; k[2] = m.z ? scale : 0.f; k[3] = m.w ? scale : 0.f;
    }
    float4 o;
    if (MODE == 0) {
      o = make_float4(x.x * sigmoidf_acc(x.x) * k[0], x.y * sigmoidf_acc(x.y) * k[1], x.z * sigmoidf_acc(x.z) * k[2], x.w * sigmoidf_acc(x.w) * k[3]);
    } else if (MODE == 1) {
      const float4 d = ldg4(b + i * 4);
      o = make_float4(d.x * k[0] * swish_grad(x.x), d.y * k[1] * swish_grad(x.y), d.z * k[2] * swish_grad(x.z), d.w * k[3] * swish_grad(x.w));
    } else if (MODE == 2) {
      const float4 r = *reinterpret_cast<const float4*>(b + i * 4);      // resid may alias out
      o = make_float4(fmaf(k[0], x.x, r.x), fmaf(k[1], x.y, r.y), fmaf(k[2], x.z, r.z), fmaf(k[3], x.w, r.w));
    } else {
      o = make_float4(k[0] * x.x, k[1] * x.y, k[2] * x.z, k[3] * x.w);
    }
    st4(out + i * 4, o);
  }
}

// ---- GLU, natural channel order: a [M, 2C] = (value | gate) -> u [M, C] = value * sigmoid(gate)  (conformer.py:30-38) ----------------
template <bool BWD>
__global__ void __launch_bounds__(256) glu_kernel(const float* __restrict__ a, const float* __restrict__ du, long long M, int C4,
                                                 float* __restrict__ out) {
  const long long total = M * C4;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long m = i / C4; const int c4 = (int)(i - m * C4);
    const float4 v = ldg4(a + (m * 2 * C4 + c4) * 4), g = ldg4(a + (m * 2 * C4 + C4 + c4) * 4);
    const float s[4] = {sigmoidf_acc(g.x), sigmoidf_acc(g.y), sigmoidf_acc(g.z), sigmoidf_acc(g.w)};
    if (!BWD) {
      st4(out + i * 4, make_float4(v.x * s[0], v.y * s[1], v.z * s[2], v.w * s[3]));
    } else {
      const float4 d = ldg4(du + i * 4);
      st4(out + (m * 2 * C4 + c4) * 4, make_float4(d.x * s[0], d.y * s[1], d.z * s[2], d.w * s[3]));
      st4(out + (m * 2 * C4 + C4 + c4) * 4, make_float4(d.x * v.x * s[0] * (1.f - s[0]), d.y * v.y * s[1] * (1.f - s[1]),
                                                        d.z * v.z * s[2] * (1.f - s[2]), d.w * v.w * s[3] * (1.f - s[3])));
    }
  }
}

// ---- finish kernels: out[c] = sum over rows of partial[row][c] (fixed order) ---------------------------------------------------------
// CTA = 32 columns x 8 row lanes: lane q sums the rows r = q (mod 8) with four independent accumulators (rows q, q + 8, q + 16, q + 24, ...), the eight
// lane sums are added in lane order through shared memory.  (One thread per column walking all rows was a serial chain of up to 592 dependent-latency
// loads in a single CTA: 25 us per call, 52 + 16 calls per training step.)  The order is fixed, so the result is deterministic.
constexpr int FIN_COLS = 32, FIN_LANES = 8;
__host__ __device__ constexpr int fin_grid(int nc) { return (nc + FIN_COLS - 1) / FIN_COLS; }
template <typename T>
__device__ __forceinline__ T fin_sum(const T* __restrict__ partial, int rows, int nc, int c, int q, T (*sm)[FIN_COLS]) {
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (c < nc) {
    int r = q;
    for (; r + 3 * FIN_LANES < rows; r += 4 * FIN_LANES) {
      a0 += partial[(long long)r * nc + c];
      a1 += partial[(long long)(r + FIN_LANES) * nc + c];
      a2 += partial[(long long)(r + 2 * FIN_LANES) * nc + c];
      a3 += partial[(long long)(r + 3 * FIN_LANES) * nc + c];
    }
    for (; r < rows; r += FIN_LANES) a0 += partial[(long long)r * nc + c];
  }
  sm[q][threadIdx.x & (FIN_COLS - 1)] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  T s = 0;
  if (q == 0) {
#pragma unroll
    for (int k = 0; k < FIN_LANES; ++k) s += sm[k][threadIdx.x & (FIN_COLS - 1)];
  }
  return s;
}
__global__ void __launch_bounds__(256) finish_f32_kernel(const float* __restrict__ partial, int rows, int nc, float* __restrict__ out0, int n0,
                                                        float* __restrict__ out1) {
  __shared__ float sm[FIN_LANES][FIN_COLS];
  const int c = blockIdx.x * FIN_COLS + (threadIdx.x & (FIN_COLS - 1)), q = threadIdx.x >> 5;
  const float s = fin_sum<float>(partial, rows, nc, c, q, sm);
  if (q == 0 && c < nc) { if (c < n0) out0[c] = s; else if (out1) out1[c - n0] = s; }
}
__global__ void __launch_bounds__(256) finish_f64_kernel(const double* __restrict__ partial, int rows, int nc, double* __restrict__ out) {
  __shared__ double sm[FIN_LANES][FIN_COLS];
  const int c = blockIdx.x * FIN_COLS + (threadIdx.x & (FIN_COLS - 1)), q = threadIdx.x >> 5;
  const double s = fin_sum<double>(partial, rows, nc, c, q, sm);
  if (q == 0 && c < nc) out[c] = s;
}

// ---- LayerNorm(64) backward: 16 lanes per token ----------------------------------------------------------------------------------------
// dx = add + rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  partial[blk][0:64] = sum dy * xhat, [64:128] = sum dy
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy,
                                                           const float* __restrict__ add, float* __restrict__ dx, long long tokens,
                                                           float* __restrict__ partial) {
  const int cq = threadIdx.x & 15, tl = threadIdx.x >> 4;
  const float4 g = ldg4(gamma + cq * 4);
  float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long tb = (long long)blockIdx.x * 16; tb < tokens; tb += (long long)gridDim.x * 16) {     // block-uniform trip count (shuffles inside)
    const long long t = tb + tl;
    const bool ok = t < tokens;
    float4 v = ok ? ldg4(x + t * 64 + cq * 4) : make_float4(0, 0, 0, 0);
    const float4 d = ok ? ldg4(dy + t * 64 + cq * 4) : make_float4(0, 0, 0, 0);
    float s = v.x + v.y + v.z + v.w;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 64.0f);
    v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
    float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / 64.0f) + 1e-5f);
    const float xh[4] = {v.x * rstd, v.y * rstd, v.z * rstd, v.w * rstd};
    const float gd[4] = {g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w};
    float m1 = gd[0] + gd[1] + gd[2] + gd[3];
    float m2 = gd[0] * xh[0] + gd[1] * xh[1] + gd[2] * xh[2] + gd[3] * xh[3];
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    m1 *= (1.0f / 64.0f); m2 *= (1.0f / 64.0f);
    if (ok) {
      float4 r = add ? *reinterpret_cast<const float4*>(add + t * 64 + cq * 4) : make_float4(0, 0, 0, 0);      // add may alias dx
      r.x += rstd * (gd[0] - m1 - xh[0] * m2); r.y += rstd * (gd[1] - m1 - xh[1] * m2);
      r.z += rstd * (gd[2] - m1 - xh[2] * m2); r.w += rstd * (gd[3] - m1 - xh[3] * m2);
      st4(dx + t * 64 + cq * 4, r);
      const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { dg[j] = fmaf(dd[j], xh[j], dg[j]); db[j] += dd[j]; }
    }
  }
  __shared__ float sh[16][128];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[tl][cq * 4 + j] = dg[j]; sh[tl][64 + cq * 4 + j] = db[j]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += sh[i][threadIdx.x];
    partial[(long long)blockIdx.x * 128 + threadIdx.x] = t;
  }
}

// ---- per-channel sums over tokens of a [M, C] matrix (C = 128): BatchNorm1d batch statistics and its backward sums --------------------
// KIND 0: (sum c, sum c^2)                                       [forward statistics, conformer.py:167 in train mode]
// KIND 1: dz = dv * swish'(scale * c + shift);  (sum dz, sum dz * chat), chat = (c - mean) * rstd
template <int KIND>
__global__ void __launch_bounds__(256) bn_sums_kernel(const float* __restrict__ c, const float* __restrict__ dv, long long M,
                                                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean_rstd,
                                                     double* __restrict__ partial) {
  constexpr int C = 128;
  const int cq = threadIdx.x & 31, tl = threadIdx.x >> 5;       // 32 lanes per token (float4 each), 8 tokens per pass
  float sc[4], sh[4], mu[4], rs[4];
  if (KIND == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { sc[j] = scale[cq * 4 + j]; sh[j] = shift[cq * 4 + j]; mu[j] = mean_rstd[cq * 4 + j]; rs[j] = mean_rstd[C + cq * 4 + j]; }
  }
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  for (long long tb = (long long)blockIdx.x * 64 + tl; tb < M; tb += (long long)gridDim.x * 64) {
    float f0[4] = {0, 0, 0, 0}, f1[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long t = tb + u * 8;
      if (t < M) {
        const float4 v = ldg4(c + t * C + cq * 4);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        if (KIND == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { f0[j] += vv[j]; f1[j] = fmaf(vv[j], vv[j], f1[j]); }
        } else {
          const float4 d = ldg4(dv + t * C + cq * 4);
          const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float dz = dd[j] * swish_grad(fmaf(vv[j], sc[j], sh[j]));
            f0[j] += dz; f1[j] = fmaf(dz, (vv[j] - mu[j]) * rs[j], f1[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s0[j] += (double)f0[j]; s1[j] += (double)f1[j]; }
  }
  __shared__ double shd[8][2 * C];
#pragma unroll
  for (int j = 0; j < 4; ++j) { shd[tl][cq * 4 + j] = s0[j]; shd[tl][C + cq * 4 + j] = s1[j]; }
  __syncthreads();
  {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += shd[i][threadIdx.x];
    partial[(long long)blockIdx.x * 2 * C + threadIdx.x] = t;
  }
}

// sums (sum c, sum c^2) over `count` tokens (after the optional cross-rank all-reduce) -> batch mean / rstd, folded scale / shift, running
// statistics (momentum update with the unbiased variance, nn.BatchNorm1d) and num_batches_tracked
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt, float momentum,
                                   float eps, float* __restrict__ scale_shift, float* __restrict__ mean_rstd) {
  constexpr int C = 128;
  const int c = threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  mean_rstd[c] = (float)mean; mean_rstd[C + c] = rstd;
  const float sc = gamma[c] * rstd;
  scale_shift[c] = sc; scale_shift[C + c] = beta[c] - (float)mean * sc;
  if (running_mean) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unb;
    if (c == 0 && nbt) *nbt += 1;
  }
}

// v = swish(scale * c + shift)
__global__ void __launch_bounds__(256) bn_swish_kernel(const float* __restrict__ c, long long M, const float* __restrict__ scale_shift, float* __restrict__ v) {
  const int cq = threadIdx.x & 31;
  const float4 sc = ldg4(scale_shift + cq * 4), sh = ldg4(scale_shift + 128 + cq * 4);
  for (long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); t < M; t += (long long)gridDim.x * 8) {
    float4 x = ldg4(c + t * 128 + cq * 4);
    x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y); x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
    st4(v + t * 128 + cq * 4, make_float4(x.x * sigmoidf_acc(x.x), x.y * sigmoidf_acc(x.y), x.z * sigmoidf_acc(x.z), x.w * sigmoidf_acc(x.w)));
  }
}

// dc = gamma * rstd * (dz - S1 / N - chat * S2 / N) with the GLOBAL sums;  block 0 also writes dgamma = S2, dbeta = S1 from the LOCAL sums
// (SyncBatchNorm returns this rank's parameter gradients; the data-parallel all-reduce of the gradients adds the ranks)
__global__ void __launch_bounds__(256) bn_swish_bwd_apply_kernel(const float* __restrict__ c, const float* __restrict__ dv, long long M,
                                                                const float* __restrict__ scale_shift, const float* __restrict__ mean_rstd,
                                                                const double* __restrict__ sums, const double* __restrict__ sums_local, double count,
                                                                float* __restrict__ dc, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int C = 128;
  const int cq = threadIdx.x & 31;
  float sc[4], sh[4], mu[4], rs[4], a1[4], a2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = cq * 4 + j;
    sc[j] = scale_shift[ch]; sh[j] = scale_shift[C + ch]; mu[j] = mean_rstd[ch]; rs[j] = mean_rstd[C + ch];
    a1[j] = (float)(sums[ch] / count); a2[j] = (float)(sums[C + ch] / count);
  }
  if (blockIdx.x == 0 && threadIdx.x < C) {
    if (dgamma) dgamma[threadIdx.x] = (float)sums_local[C + threadIdx.x];
    if (dbeta) dbeta[threadIdx.x] = (float)sums_local[threadIdx.x];
  }
  for (long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); t < M; t += (long long)gridDim.x * 8) {
    const float4 x = ldg4(c + t * C + cq * 4), d = ldg4(dv + t * C + cq * 4);
    const float xx[4] = {x.x, x.y, x.z, x.w}, dd[4] = {d.x, d.y, d.z, d.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dz = dd[j] * swish_grad(fmaf(xx[j], sc[j], sh[j]));
      const float ch = (xx[j] - mu[j]) * rs[j];
      o[j] = sc[j] * (dz - a1[j] - ch * a2[j]);          // sc = gamma * rstd
    }
    st4(dc + t * C + cq * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ---- depthwise conv weight gradient: dw[ch][k] = sum dc[pos] * u[pos + k - 15], db[ch] = sum dc (conformer.py:40-48) -----------------
constexpr int DWG_TI = 32, DWG_K = 31, DWG_PAD = 15, DWG_C = 128;
__global__ void __launch_bounds__(128) dwconv_wgrad_kernel(const float* __restrict__ u, const float* __restrict__ dc, const SebSeq sq, int nchunks,
                                                          long long nitems, float* __restrict__ partial) {
  __shared__ __align__(16) float ut[DWG_TI + DWG_K - 1][DWG_C];
  __shared__ __align__(16) float dt[DWG_TI][DWG_C];
  const int ch = threadIdx.x;
  float acc[DWG_K];
#pragma unroll
  for (int k = 0; k < DWG_K; ++k) acc[k] = 0.f;
  float bacc = 0.f;
  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int seq = (int)(item / nchunks), i0 = (int)(item - (long long)seq * nchunks) * DWG_TI;
    const long long base = (long long)(seq / sq.inner) * sq.outer_stride + (seq % sq.inner);
    __syncthreads();
    for (int idx = threadIdx.x; idx < (DWG_TI + DWG_K - 1) * (DWG_C / 4); idx += 128) {
      const int r = idx >> 5, c4 = idx & 31;
      const int i = i0 + r - DWG_PAD;
      float4 v = make_float4(0, 0, 0, 0);
      if (i >= 0 && i < sq.n) v = ldg4(u + (base + (long long)i * sq.pos_stride) * DWG_C + c4 * 4);
      *reinterpret_cast<float4*>(&ut[r][c4 * 4]) = v;
    }
    for (int idx = threadIdx.x; idx < DWG_TI * (DWG_C / 4); idx += 128) {
      const int r = idx >> 5, c4 = idx & 31;
      const int i = i0 + r;
      float4 v = make_float4(0, 0, 0, 0);
      if (i < sq.n) v = ldg4(dc + (base + (long long)i * sq.pos_stride) * DWG_C + c4 * 4);
      *reinterpret_cast<float4*>(&dt[r][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int p = 0; p < DWG_TI; ++p) {
      const float d = dt[p][ch];
      bacc += d;
#pragma unroll
      for (int k = 0; k < DWG_K; ++k) acc[k] = fmaf(d, ut[p + k][ch], acc[k]);
    }
  }
  float* row = partial + (long long)blockIdx.x * (DWG_C * 32);
#pragma unroll
  for (int k = 0; k < DWG_K; ++k) row[ch * DWG_K + k] = acc[k];        // [ch][k]: the parameter's own layout (128, 1, 31)
  row[DWG_C * DWG_K + ch] = bacc;
}

// ---- InstanceNorm2d(affine) + PReLU backward (generator.py:21-22,40-41,46-47,101-102,120-121) ---------------------------------------------
// z = gamma * xhat + beta, y = z >= 0 ? z : slope * z.   dz = dy * (z >= 0 ? 1 : slope)
// sums[b][c][0..2] = (sum dz, sum dz * xhat, sum dy * min(z, 0)) over the (T, F) plane, fp64 two-stage
constexpr int INB_ROWS = 1024;
__global__ void __launch_bounds__(256) inorm_prelu_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long pix_per_b,
                                                                  const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ slope, double* __restrict__ part) {
  constexpr int C = 64;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int cq = threadIdx.x & 15, pl = threadIdx.x >> 4;
  float mu[4], rs[4], ga[4], be[4], sl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cq * 4 + j;
    mu[j] = stats[((long long)b * C + c) * 2]; rs[j] = stats[((long long)b * C + c) * 2 + 1];
    ga[j] = gamma[c]; be[j] = beta[c]; sl[j] = slope[c];
  }
  const long long p0 = (long long)chunk * INB_ROWS;
  long long p1 = p0 + INB_ROWS; if (p1 > pix_per_b) p1 = pix_per_b;
  const float* xb = x + (long long)b * pix_per_b * C;
  const float* db = dy + (long long)b * pix_per_b * C;
  double s[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  for (long long p = p0 + pl; p < p1; p += 16 * 8) {
    float f[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int uu = 0; uu < 8; ++uu) {
      const long long pp = p + uu * 16;
      if (pp < p1) {
        const float4 v = ldg4(xb + pp * C + cq * 4), d = ldg4(db + pp * C + cq * 4);
        const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (vv[j] - mu[j]) * rs[j];
          const float z = fmaf(ga[j], xh, be[j]);
          const float dz = z >= 0.f ? dd[j] : dd[j] * sl[j];
          f[0][j] += dz; f[1][j] = fmaf(dz, xh, f[1][j]); f[2][j] += z >= 0.f ? 0.f : dd[j] * z;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[k][j] += (double)f[k][j];
  }
  __shared__ double sh[16][C * 3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) sh[pl][(cq * 4 + j) * 3 + k] = s[k][j];
  __syncthreads();
  if (threadIdx.x < C * 3) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += sh[i][threadIdx.x];
    part[((long long)b * gridDim.x + chunk) * (C * 3) + threadIdx.x] = t;
  }
}

// C = 1 (MaskDecoder.norm / .prelu): one value per pixel
__global__ void __launch_bounds__(256) inorm_prelu_bwd_sums1_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long pix_per_b,
                                                                   const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const float* __restrict__ slope, double* __restrict__ part) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float mu = stats[b * 2], rs = stats[b * 2 + 1], ga = gamma[0], be = beta[0], sl = slope[0];
  const long long p0 = (long long)chunk * INB_ROWS * 64;
  long long p1 = p0 + (long long)INB_ROWS * 64; if (p1 > pix_per_b) p1 = pix_per_b;
  const float* xb = x + (long long)b * pix_per_b;
  const float* db = dy + (long long)b * pix_per_b;
  double s[3] = {0, 0, 0};
  for (long long p = p0 + threadIdx.x; p < p1; p += 256) {
    const float xh = (xb[p] - mu) * rs, d = db[p];
    const float z = fmaf(ga, xh, be);
    const float dz = z >= 0.f ? d : d * sl;
    s[0] += (double)dz; s[1] += (double)(dz * xh); s[2] += z >= 0.f ? 0.0 : (double)(d * z);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  __shared__ double sh[8][3];
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = s[0]; sh[threadIdx.x >> 5][1] = s[1]; sh[threadIdx.x >> 5][2] = s[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    part[((long long)b * gridDim.x + chunk) * 3 + threadIdx.x] = t;
  }
}

// sums[b][c][k] = sum over chunks; dgamma[c] = sum_b S2, dbeta[c] = sum_b S1, dslope[c] = sum_b S3  (one CTA, C * 3 threads)
__global__ void inorm_prelu_bwd_finish_kernel(const double* __restrict__ part, int B, int chunks, int C, double* __restrict__ sums,
                                              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dslope) {
  const int i = threadIdx.x;
  if (i >= C * 3) return;
  double tot = 0.0;
  for (int b = 0; b < B; ++b) {
    double s = 0.0;
    for (int k = 0; k < chunks; ++k) s += part[((long long)b * chunks + k) * (C * 3) + i];
    sums[(long long)b * C * 3 + i] = s;
    tot += s;
  }
  const int c = i / 3, k = i - c * 3;
  float* dst = k == 0 ? dbeta : (k == 1 ? dgamma : dslope);
  if (dst) dst[c] = (float)tot;
}

// dx = gamma * rstd * (dz - S1 / N - xhat * S2 / N)
template <int C>
__global__ void __launch_bounds__(256) inorm_prelu_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long pix_per_b,
                                                                   const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const float* __restrict__ slope,
                                                                   const double* __restrict__ sums, float* __restrict__ dx) {
  const int b = blockIdx.y;
  const double inv_n = 1.0 / (double)pix_per_b;
  if (C == 64) {
    const int cq = threadIdx.x & 15;
    float mu[4], rs[4], ga[4], be[4], sl[4], a1[4], a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cq * 4 + j;
      mu[j] = stats[((long long)b * 64 + c) * 2]; rs[j] = stats[((long long)b * 64 + c) * 2 + 1];
      ga[j] = gamma[c]; be[j] = beta[c]; sl[j] = slope[c];
      a1[j] = (float)(sums[((long long)b * 64 + c) * 3] * inv_n); a2[j] = (float)(sums[((long long)b * 64 + c) * 3 + 1] * inv_n);
    }
    const float* xb = x + (long long)b * pix_per_b * 64;
    const float* db = dy + (long long)b * pix_per_b * 64;
    float* ob = dx + (long long)b * pix_per_b * 64;
    for (long long p = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); p < pix_per_b; p += (long long)gridDim.x * 16) {
      const float4 v = ldg4(xb + p * 64 + cq * 4), d = ldg4(db + p * 64 + cq * 4);
      const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (vv[j] - mu[j]) * rs[j];
        const float z = fmaf(ga[j], xh, be[j]);
        const float dz = z >= 0.f ? dd[j] : dd[j] * sl[j];
        o[j] = ga[j] * rs[j] * (dz - a1[j] - xh * a2[j]);
      }
      st4(ob + p * 64 + cq * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
  } else {
    const float mu = stats[b * 2], rs = stats[b * 2 + 1], ga = gamma[0], be = beta[0], sl = slope[0];
    const float a1 = (float)(sums[(long long)b * 3] * inv_n), a2 = (float)(sums[(long long)b * 3 + 1] * inv_n);
    const float* xb = x + (long long)b * pix_per_b;
    const float* db = dy + (long long)b * pix_per_b;
    float* ob = dx + (long long)b * pix_per_b;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < pix_per_b; p += (long long)gridDim.x * 256) {
      const float xh = (xb[p] - mu) * rs, d = db[p];
      const float z = fmaf(ga, xh, be);
      const float dz = z >= 0.f ? d : d * sl;
      ob[p] = ga * rs * (dz - a1 - xh * a2);
    }
  }
}

// ---- decoder heads: Conv2d(64 -> NO, (1, 2)) on x [rows, Fin, 64]; weights in the PARAMETER layout (NO, 64, 1, 2) -------------------------
template <int NO>
__global__ void __launch_bounds__(256) head_conv_kernel(const float* __restrict__ x, long long rows, int Fin, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out) {
  const int cq = threadIdx.x & 15;
  float w0[NO][4], w1[NO][4];
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int j = 0; j < 4; ++j) { w0[o][j] = w[(o * 64 + cq * 4 + j) * 2]; w1[o][j] = w[(o * 64 + cq * 4 + j) * 2 + 1]; }
  const int Fo = Fin - 1;
  const long long total = rows * Fo;
  for (long long ob = (long long)blockIdx.x * 16; ob < total; ob += (long long)gridDim.x * 16) {
    const long long o = ob + (threadIdx.x >> 4);
    const bool ok = o < total;
    float acc[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) acc[k] = 0.f;
    if (ok) {
      const long long r = o / Fo; const int f = (int)(o - r * Fo);
      const float* p = x + (r * Fin + f) * 64 + cq * 4;
      const float4 a = ldg4(p), b = ldg4(p + 64);
#pragma unroll
      for (int k = 0; k < NO; ++k)
        acc[k] = a.x * w0[k][0] + a.y * w0[k][1] + a.z * w0[k][2] + a.w * w0[k][3] + b.x * w1[k][0] + b.y * w1[k][1] + b.z * w1[k][2] + b.w * w1[k][3];
    }
#pragma unroll
    for (int k = 0; k < NO; ++k)
#pragma unroll
      for (int s = 1; s < 16; s <<= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
    if (ok && cq == 0) {
#pragma unroll
      for (int k = 0; k < NO; ++k) out[o * NO + k] = acc[k] + bias[k];
    }
  }
}

// dx[r, f, c] = sum_o dout[r, f, o] w[o][c][0] (f < Fo) + dout[r, f - 1, o] w[o][c][1] (f >= 1)
// partial[blk] = (dw [NO][64][2] | db [NO]):  dw[o][c][tap] = sum dout[r, f - tap, o] x[r, f, c]
template <int NO>
__global__ void __launch_bounds__(256) head_conv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, long long rows, int Fin,
                                                           const float* __restrict__ w, float* __restrict__ dx, float* __restrict__ partial) {
  const int cq = threadIdx.x & 15, pl = threadIdx.x >> 4;
  float w0[NO][4], w1[NO][4];
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int j = 0; j < 4; ++j) { w0[o][j] = w[(o * 64 + cq * 4 + j) * 2]; w1[o][j] = w[(o * 64 + cq * 4 + j) * 2 + 1]; }
  const int Fo = Fin - 1;
  const long long total = rows * Fin;
  float g0[NO][4], g1[NO][4], gb[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) { gb[o] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { g0[o][j] = 0.f; g1[o][j] = 0.f; } }
  for (long long i = (long long)blockIdx.x * 16 + pl; i < total; i += (long long)gridDim.x * 16) {
    const long long r = i / Fin; const int f = (int)(i - r * Fin);
    const float4 v = ldg4(x + i * 64 + cq * 4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    float d0[NO], d1[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      d0[o] = f < Fo ? dout[(r * Fo + f) * NO + o] : 0.f;
      d1[o] = f >= 1 ? dout[(r * Fo + f - 1) * NO + o] : 0.f;
      if (cq == 0) gb[o] += d0[o];
    }
    float o4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o4[j] = fmaf(d0[o], w0[o][j], fmaf(d1[o], w1[o][j], o4[j]));
        g0[o][j] = fmaf(d0[o], vv[j], g0[o][j]);
        g1[o][j] = fmaf(d1[o], vv[j], g1[o][j]);
      }
    st4(dx + i * 64 + cq * 4, make_float4(o4[0], o4[1], o4[2], o4[3]));
  }
  constexpr int NC = NO * 128 + NO;
  __shared__ float sh[16][NO * 128 + 4];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { sh[pl][(o * 64 + cq * 4 + j) * 2] = g0[o][j]; sh[pl][(o * 64 + cq * 4 + j) * 2 + 1] = g1[o][j]; }
    if (cq == 0) sh[pl][NO * 128 + o] = gb[o];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NC; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += sh[i][c];
    partial[(long long)blockIdx.x * NC + c] = t;
  }
}

// ---- mask tail + recombination with DEVICE scalars (parameters change every step: no host read) --------------------------------------
// scal[0..4] = pointers to norm.weight, norm.bias, prelu.weight, final_conv.weight, final_conv.bias (one float each)
struct MaskScal { const float* p[5]; };
__global__ void __launch_bounds__(256) mask_recombine_dev_kernel(const float* __restrict__ raw, const float* __restrict__ stats, long long rows_per_b, int F,
                                                                const MaskScal sc5, const float* __restrict__ slope_f, const float* __restrict__ in3,
                                                                const float* __restrict__ cplx, float* __restrict__ est) {
  const int b = blockIdx.y;
  const float in_gamma = *sc5.p[0], in_beta = *sc5.p[1], slope1 = *sc5.p[2], wf = *sc5.p[3], bf = *sc5.p[4];
  const float mean = stats[b * 2], rstd = stats[b * 2 + 1];
  const float sc = rstd * in_gamma, sh = in_beta - mean * sc;
  const long long total = rows_per_b * F, base = (long long)b * total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    float v = fmaf(raw[base + i], sc, sh);
    v = v >= 0.f ? v : v * slope1;
    v = fmaf(v, wf, bf);
    const float sf = slope_f[f];
    v = v >= 0.f ? v : v * sf;
    const float re = in3[(base + i) * 3 + 1], im = in3[(base + i) * 3 + 2];
    const float2 c = *reinterpret_cast<const float2*>(cplx + (base + i) * 2);
    *reinterpret_cast<float2*>(est + (base + i) * 2) = make_float2(fmaf(v, re, c.x), fmaf(v, im, c.y));
  }
}

// backward of  est = PReLU_f(wf * p1 + bf) * (re, im) + cplx  w.r.t. p1 = PReLU(IN(raw)) and the tail's parameters.
// grid (chunks, B), thread = frequency bin (blockDim = 256 >= F); partial[b * chunks + chunk] = (dslope_f [F] | dwf | dbf)
__global__ void __launch_bounds__(256) mask_tail_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ stats, long long rows_per_b, int F,
                                                           int rows_per_chunk, const MaskScal sc5, const float* __restrict__ slope_f,
                                                           const float* __restrict__ in3, const float* __restrict__ dest, float* __restrict__ dp1,
                                                           float* __restrict__ partial) {
  const int b = blockIdx.y, f = threadIdx.x;
  const float in_gamma = *sc5.p[0], in_beta = *sc5.p[1], slope1 = *sc5.p[2], wf = *sc5.p[3], bf = *sc5.p[4];
  const float mean = stats[b * 2], rstd = stats[b * 2 + 1];
  const float sc = rstd * in_gamma, sh = in_beta - mean * sc;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk; if (r1 > rows_per_b) r1 = rows_per_b;
  float ds = 0.f, dwf = 0.f, dbf = 0.f;
  if (f < F) {
    const float sf = slope_f[f];
    for (long long r = r0; r < r1; ++r) {
      const long long i = ((long long)b * rows_per_b + r) * F + f;
      float z1 = fmaf(raw[i], sc, sh);
      const float p1 = z1 >= 0.f ? z1 : z1 * slope1;
      const float m2 = fmaf(p1, wf, bf);
      const float2 d = *reinterpret_cast<const float2*>(dest + i * 2);
      const float dmask = d.x * in3[i * 3 + 1] + d.y * in3[i * 3 + 2];
      const float dm2 = m2 >= 0.f ? dmask : dmask * sf;
      ds += m2 >= 0.f ? 0.f : dmask * m2;
      dwf = fmaf(dm2, p1, dwf); dbf += dm2;
      dp1[i] = dm2 * wf;
    }
  }
  __shared__ float sh2[2][8];
  float a = dwf, c = dbf;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if ((threadIdx.x & 31) == 0) { sh2[0][threadIdx.x >> 5] = a; sh2[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  float* row = partial + ((long long)b * gridDim.x + blockIdx.x) * (F + 2);
  if (f < F) row[f] = ds;
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh2[threadIdx.x][i];
    row[F + threadIdx.x] = t;
  }
}

// ---- conv_1[0] (1x1, 3 -> 64) weight gradient: dw[c][k] = sum g[p][c] in3[p][k], db[c] = sum g[p][c] -------------------------------------
__global__ void __launch_bounds__(256) conv1x1_in3_wgrad_kernel(const float* __restrict__ in3, const float* __restrict__ g, long long pixels,
                                                               float* __restrict__ partial) {
  const int cq = threadIdx.x & 15, pl = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
  for (long long p = (long long)blockIdx.x * 16 + pl; p < pixels; p += (long long)gridDim.x * 16) {
    const float4 d = ldg4(g + p * 64 + cq * 4);
    const float a = in3[p * 3], b = in3[p * 3 + 1], c = in3[p * 3 + 2];
    const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j][0] = fmaf(dd[j], a, acc[j][0]); acc[j][1] = fmaf(dd[j], b, acc[j][1]); acc[j][2] = fmaf(dd[j], c, acc[j][2]); acc[j][3] += dd[j];
    }
  }
  __shared__ float sh[16][256];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cq * 4 + j;
    sh[pl][c * 3] = acc[j][0]; sh[pl][c * 3 + 1] = acc[j][1]; sh[pl][c * 3 + 2] = acc[j][2]; sh[pl][192 + c] = acc[j][3];
  }
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += sh[i][threadIdx.x];
  partial[(long long)blockIdx.x * 256 + threadIdx.x] = t;
}

__global__ void __launch_bounds__(256) merge_ri_kernel(const float* __restrict__ re, const float* __restrict__ im, long long n, float2* __restrict__ est) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) est[i] = make_float2(re[i], im[i]);
}

// fp32 [M, 192] (q | k | v) -> fp16 [M, 192] with q scaled by dim_head^-0.5 * log2(e): the tensor-core attention's input format
__global__ void __launch_bounds__(256) qkv_to_f16_kernel(const float* __restrict__ qkv, long long M, __half* __restrict__ out) {
  const long long total = M * 48;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c4 = (int)(i % 48);
    const float4 v = ldg4(qkv + i * 4);
    const float sc = c4 < 16 ? 0.25f * 1.4426950408889634f : 1.0f;
    const __half2 a = __floats2half2_rn(v.x * sc, v.y * sc), b = __floats2half2_rn(v.z * sc, v.w * sc);
    *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

}  // namespace seb

using namespace seb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int seb200_dropout_mask(unsigned char* mask, long long n, float p, unsigned long long seed, unsigned long long offset, void* stream) {
  SEB_REQUIRE(mask && n > 0 && p >= 0.f && p < 1.f && (reinterpret_cast<uintptr_t>(mask) & 3u) == 0, SEB_EINVAL, "dropout_mask: bad arguments");
  const long long n4 = (n + 3) / 4;
  dropout_mask_kernel<<<tgrid(n4, 256 * 4), 256, 0, ST(stream)>>>(mask, n4, n, p, seed, offset);
  SEB_CHECK_LAUNCH("dropout_mask_kernel");
  return 0;
}

static int elem_check(const void* a, const void* out, long long n, const char* what) {
  SEB_REQUIRE(a && out && n > 0 && n % 4 == 0 && aligned16(a) && aligned16(out), SEB_EINVAL, "%s: bad arguments (n must be a multiple of 4, pointers 16-byte aligned)", what);
  return 0;
}

extern "C" int seb200_swish_dropout(const float* a, const unsigned char* mask, float scale, float* h, long long n, void* stream) {
  if (int rc = elem_check(a, h, n, "swish_dropout")) return rc;
  elem_mask_kernel<0><<<tgrid(n / 4, 256 * 4), 256, 0, ST(stream)>>>(a, nullptr, mask, scale, h, n / 4);
  SEB_CHECK_LAUNCH("elem_mask_kernel<swish_dropout>");
  return 0;
}
extern "C" int seb200_swish_dropout_bwd(const float* a, const unsigned char* mask, float scale, const float* dh, float* da, long long n, void* stream) {
  if (int rc = elem_check(a, da, n, "swish_dropout_bwd")) return rc;
  SEB_REQUIRE(dh && aligned16(dh), SEB_EINVAL, "swish_dropout_bwd: dh null / unaligned");
  elem_mask_kernel<1><<<tgrid(n / 4, 256 * 4), 256, 0, ST(stream)>>>(a, dh, mask, scale, da, n / 4);
  SEB_CHECK_LAUNCH("elem_mask_kernel<swish_dropout_bwd>");
  return 0;
}
extern "C" int seb200_dropout_residual(const float* t, const unsigned char* mask, float scale, const float* resid, float* y, long long n, void* stream) {
  if (int rc = elem_check(t, y, n, "dropout_residual")) return rc;
  SEB_REQUIRE(resid && aligned16(resid), SEB_EINVAL, "dropout_residual: resid null / unaligned");
  elem_mask_kernel<2><<<tgrid(n / 4, 256 * 4), 256, 0, ST(stream)>>>(t, resid, mask, scale, y, n / 4);
  SEB_CHECK_LAUNCH("elem_mask_kernel<dropout_residual>");
  return 0;
}
extern "C" int seb200_scale_mask(const float* dy, const unsigned char* mask, float scale, float* dt, long long n, void* stream) {
  if (int rc = elem_check(dy, dt, n, "scale_mask")) return rc;
  elem_mask_kernel<3><<<tgrid(n / 4, 256 * 4), 256, 0, ST(stream)>>>(dy, nullptr, mask, scale, dt, n / 4);
  SEB_CHECK_LAUNCH("elem_mask_kernel<scale_mask>");
  return 0;
}

extern "C" int seb200_glu(const float* a, long long M, int C, float* u, void* stream) {
  SEB_REQUIRE(a && u && M > 0 && C > 0 && C % 4 == 0 && aligned16(a) && aligned16(u), SEB_EINVAL, "glu: bad arguments");
  glu_kernel<false><<<tgrid(M * (C / 4), 256 * 4), 256, 0, ST(stream)>>>(a, nullptr, M, C / 4, u);
  SEB_CHECK_LAUNCH("glu_kernel");
  return 0;
}
extern "C" int seb200_glu_bwd(const float* a, const float* du, long long M, int C, float* da, void* stream) {
  SEB_REQUIRE(a && du && da && M > 0 && C > 0 && C % 4 == 0 && aligned16(a) && aligned16(du) && aligned16(da), SEB_EINVAL, "glu_bwd: bad arguments");
  glu_kernel<true><<<tgrid(M * (C / 4), 256 * 4), 256, 0, ST(stream)>>>(a, du, M, C / 4, da);
  SEB_CHECK_LAUNCH("glu_kernel<bwd>");
  return 0;
}

// LayerNorm(64) backward.  dx = (add ? add : 0) + dLN(x; dy) (add / dx may alias); dgamma / dbeta [64].  workspace >= seb200_train_workspace_floats()
constexpr int TRAIN_WS_FLOATS = 148 * 8 * 512;
extern "C" long long seb200_train_workspace_floats(void) { return TRAIN_WS_FLOATS; }

extern "C" int seb200_layernorm_bwd(const float* x, const float* gamma, const float* dy, const float* add, float* dx, long long tokens,
                                    float* dgamma, float* dbeta, float* workspace, void* stream) {
  SEB_REQUIRE(x && gamma && dy && dx && dgamma && dbeta && workspace && tokens > 0 && aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma), SEB_EINVAL,
              "layernorm_bwd: bad arguments");
  const int nb = tgrid(tokens, 16 * 16, 148 * 4);
  layernorm_bwd_kernel<<<nb, 256, 0, ST(stream)>>>(x, gamma, dy, add, dx, tokens, workspace);
  SEB_CHECK_LAUNCH("layernorm_bwd_kernel");
  finish_f32_kernel<<<fin_grid(128), 256, 0, ST(stream)>>>(workspace, nb, 128, dgamma, 64, dbeta);
  SEB_CHECK_LAUNCH("finish_f32_kernel");
  return 0;
}

// BatchNorm1d(128) in train mode, step 1: LOCAL sums[0:128] = sum c, sums[128:256] = sum c^2 over the M tokens (fp64).  workspace: doubles.
extern "C" int seb200_bn_sums(const float* c, long long M, double* sums, double* workspace, void* stream) {
  SEB_REQUIRE(c && sums && workspace && M > 0 && aligned16(c), SEB_EINVAL, "bn_sums: bad arguments");
  const int nb = tgrid(M, 64 * 4, 148 * 4);
  bn_sums_kernel<0><<<nb, 256, 0, ST(stream)>>>(c, nullptr, M, nullptr, nullptr, nullptr, workspace);
  SEB_CHECK_LAUNCH("bn_sums_kernel");
  finish_f64_kernel<<<fin_grid(256), 256, 0, ST(stream)>>>(workspace, nb, 256, sums);
  SEB_CHECK_LAUNCH("finish_f64_kernel");
  return 0;
}
// step 2 (after the optional all-reduce of sums / count across ranks): batch mean / rstd, folded scale / shift, running statistics
extern "C" int seb200_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                  long long* num_batches_tracked, float momentum, float eps, float* scale_shift, float* mean_rstd, void* stream) {
  SEB_REQUIRE(sums && gamma && beta && scale_shift && mean_rstd && count >= 1.0, SEB_EINVAL, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<1, 128, 0, ST(stream)>>>(sums, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, scale_shift, mean_rstd);
  SEB_CHECK_LAUNCH("bn_finalize_kernel");
  return 0;
}
extern "C" int seb200_bn_swish(const float* c, long long M, const float* scale_shift, float* v, void* stream) {
  SEB_REQUIRE(c && scale_shift && v && M > 0 && aligned16(c) && aligned16(v) && aligned16(scale_shift), SEB_EINVAL, "bn_swish: bad arguments");
  bn_swish_kernel<<<tgrid(M, 8 * 8), 256, 0, ST(stream)>>>(c, M, scale_shift, v);
  SEB_CHECK_LAUNCH("bn_swish_kernel");
  return 0;
}
// backward step 1: LOCAL sums[0:128] = sum dz, sums[128:256] = sum dz * chat
extern "C" int seb200_bn_swish_bwd_sums(const float* c, const float* dv, long long M, const float* scale_shift, const float* mean_rstd, double* sums,
                                        double* workspace, void* stream) {
  SEB_REQUIRE(c && dv && scale_shift && mean_rstd && sums && workspace && M > 0 && aligned16(c) && aligned16(dv), SEB_EINVAL, "bn_swish_bwd_sums: bad arguments");
  const int nb = tgrid(M, 64 * 4, 148 * 4);
  bn_sums_kernel<1><<<nb, 256, 0, ST(stream)>>>(c, dv, M, scale_shift, scale_shift + 128, mean_rstd, workspace);
  SEB_CHECK_LAUNCH("bn_sums_kernel<bwd>");
  finish_f64_kernel<<<fin_grid(256), 256, 0, ST(stream)>>>(workspace, nb, 256, sums);
  SEB_CHECK_LAUNCH("finish_f64_kernel");
  return 0;
}
// backward step 2 (after the optional all-reduce): dc from the global sums; dgamma = S2, dbeta = S1 of `sums_local` (this rank's own sums; NULL = sums)
extern "C" int seb200_bn_swish_bwd_apply(const float* c, const float* dv, long long M, const float* scale_shift, const float* mean_rstd, const double* sums,
                                         const double* sums_local, double count, float* dc, float* dgamma, float* dbeta, void* stream) {
  SEB_REQUIRE(c && dv && scale_shift && mean_rstd && sums && dc && M > 0 && count >= 1.0 && aligned16(c) && aligned16(dv) && aligned16(dc), SEB_EINVAL,
              "bn_swish_bwd_apply: bad arguments");
  bn_swish_bwd_apply_kernel<<<tgrid(M, 8 * 8), 256, 0, ST(stream)>>>(c, dv, M, scale_shift, mean_rstd, sums, sums_local ? sums_local : sums, count, dc, dgamma, dbeta);
  SEB_CHECK_LAUNCH("bn_swish_bwd_apply_kernel");
  return 0;
}

// depthwise conv weight gradient: dw in the parameter layout (128, 1, 31), db [128]
extern "C" int seb200_dwconv_wgrad(const float* u, const float* dc, const SebSeq* seq, float* dw, float* db, float* workspace, void* stream) {
  SEB_REQUIRE(u && dc && seq && dw && db && workspace && aligned16(u) && aligned16(dc), SEB_EINVAL, "dwconv_wgrad: bad arguments");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0, SEB_EINVAL, "dwconv_wgrad: bad sequence descriptor");
  const int nchunks = (seq->n + DWG_TI - 1) / DWG_TI;
  const long long nitems = (long long)seq->nseq * nchunks;
  const int nb = (int)(nitems < 148 ? nitems : 148);
  dwconv_wgrad_kernel<<<nb, 128, 0, ST(stream)>>>(u, dc, *seq, nchunks, nitems, workspace);
  SEB_CHECK_LAUNCH("dwconv_wgrad_kernel");
  finish_f32_kernel<<<fin_grid(DWG_C * 32), 256, 0, ST(stream)>>>(workspace, nb, DWG_C * 32, dw, DWG_C * DWG_K, db);
  SEB_CHECK_LAUNCH("finish_f32_kernel");
  return 0;
}

// InstanceNorm2d(affine) + PReLU backward, C = 64 or 1.  sums: doubles [B][C][3]; workspace: doubles, >= B * chunks * C * 3
extern "C" long long seb200_inorm_bwd_workspace_doubles(int B, long long pix_per_b, int C) {
  const long long per = (C == 1) ? (long long)INB_ROWS * 64 : INB_ROWS;
  return (long long)B * ((pix_per_b + per - 1) / per) * C * 3 + (long long)B * C * 3;
}
extern "C" int seb200_inorm_prelu_bwd(const float* x, const float* dy, int B, long long pix_per_b, int C, const float* stats, const float* gamma,
                                      const float* beta, const float* slope, float* dx, float* dgamma, float* dbeta, float* dslope,
                                      double* workspace, long long workspace_doubles, void* stream) {
  SEB_REQUIRE(x && dy && stats && gamma && beta && slope && dx && workspace && B > 0 && B < 65536 && pix_per_b > 0 && (C == 64 || C == 1), SEB_EINVAL,
              "inorm_prelu_bwd: bad arguments (C must be 64 or 1)");
  SEB_REQUIRE(workspace_doubles >= seb200_inorm_bwd_workspace_doubles(B, pix_per_b, C), SEB_EINVAL, "inorm_prelu_bwd: workspace too small");
  if (C == 64) SEB_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx), SEB_EALIGN, "inorm_prelu_bwd: unaligned tensor");
  const long long per = (C == 1) ? (long long)INB_ROWS * 64 : INB_ROWS;
  const int chunks = (int)((pix_per_b + per - 1) / per);
  double* sums = workspace;
  double* part = workspace + (long long)B * C * 3;
  dim3 grid(chunks, B);
  if (C == 64) inorm_prelu_bwd_sums_kernel<<<grid, 256, 0, ST(stream)>>>(x, dy, pix_per_b, stats, gamma, beta, slope, part);
  else inorm_prelu_bwd_sums1_kernel<<<grid, 256, 0, ST(stream)>>>(x, dy, pix_per_b, stats, gamma, beta, slope, part);
  SEB_CHECK_LAUNCH("inorm_prelu_bwd_sums_kernel");
  inorm_prelu_bwd_finish_kernel<<<1, 192, 0, ST(stream)>>>(part, B, chunks, C, sums, dgamma, dbeta, dslope);
  SEB_CHECK_LAUNCH("inorm_prelu_bwd_finish_kernel");
  dim3 g2(tgrid(pix_per_b, (C == 64 ? 16 : 256) * 8, 148 * 8 / (B < 8 ? B : 8) + 1), B);
  if (C == 64) inorm_prelu_bwd_apply_kernel<64><<<g2, 256, 0, ST(stream)>>>(x, dy, pix_per_b, stats, gamma, beta, slope, sums, dx);
  else inorm_prelu_bwd_apply_kernel<1><<<g2, 256, 0, ST(stream)>>>(x, dy, pix_per_b, stats, gamma, beta, slope, sums, dx);
  SEB_CHECK_LAUNCH("inorm_prelu_bwd_apply_kernel");
  return 0;
}

// decoder heads with parameter-layout weights (NO, 64, 1, 2): forward and backward
extern "C" int seb200_head_conv(const float* x, long long rows, int Fin, const float* w, const float* bias, int NO, float* out, void* stream) {
  SEB_REQUIRE(x && w && bias && out && rows > 0 && Fin > 1 && (NO == 1 || NO == 2) && aligned16(x), SEB_EINVAL, "head_conv: bad arguments");
  const int nb = tgrid(rows * (Fin - 1), 16 * 8);
  if (NO == 1) head_conv_kernel<1><<<nb, 256, 0, ST(stream)>>>(x, rows, Fin, w, bias, out);
  else head_conv_kernel<2><<<nb, 256, 0, ST(stream)>>>(x, rows, Fin, w, bias, out);
  SEB_CHECK_LAUNCH("head_conv_kernel");
  return 0;
}
extern "C" int seb200_head_conv_bwd(const float* x, const float* dout, long long rows, int Fin, const float* w, int NO, float* dx, float* dw, float* db,
                                    float* workspace, void* stream) {
  SEB_REQUIRE(x && dout && w && dx && dw && db && workspace && rows > 0 && Fin > 1 && (NO == 1 || NO == 2) && aligned16(x) && aligned16(dx), SEB_EINVAL,
              "head_conv_bwd: bad arguments");
  const int nb = tgrid(rows * Fin, 16 * 16, 148 * 4);
  const int nc = NO * 128 + NO;
  if (NO == 1) head_conv_bwd_kernel<1><<<nb, 256, 0, ST(stream)>>>(x, dout, rows, Fin, w, dx, workspace);
  else head_conv_bwd_kernel<2><<<nb, 256, 0, ST(stream)>>>(x, dout, rows, Fin, w, dx, workspace);
  SEB_CHECK_LAUNCH("head_conv_bwd_kernel");
  finish_f32_kernel<<<fin_grid(nc), 256, 0, ST(stream)>>>(workspace, nb, nc, dw, NO * 128, db);
  SEB_CHECK_LAUNCH("finish_f32_kernel");
  return 0;
}

// mask tail + recombination, scalars by device pointer (norm.weight, norm.bias, prelu.weight, final_conv.weight, final_conv.bias)
extern "C" int seb200_mask_recombine_dev(const float* mask_raw, const float* mask_stats, int B, long long rows_per_b, int F, const float* const* scalars5,
                                         const float* slope_f, const float* in3, const float* cplx, float* est, void* stream) {
  SEB_REQUIRE(mask_raw && mask_stats && scalars5 && slope_f && in3 && cplx && est && B > 0 && B < 65536 && rows_per_b > 0 && F > 0, SEB_EINVAL, "mask_recombine_dev: bad arguments");
  MaskScal ms;
  for (int i = 0; i < 5; ++i) { SEB_REQUIRE(scalars5[i], SEB_EINVAL, "mask_recombine_dev: null scalar %d", i); ms.p[i] = scalars5[i]; }
  dim3 grid(tgrid(rows_per_b * F, 256 * 4, 148 * 8 / (B < 8 ? B : 8) + 1), B);
  mask_recombine_dev_kernel<<<grid, 256, 0, ST(stream)>>>(mask_raw, mask_stats, rows_per_b, F, ms, slope_f, in3, cplx, est);
  SEB_CHECK_LAUNCH("mask_recombine_dev_kernel");
  return 0;
}
// backward: d_est [B*T, F, 2] -> dp1 [B*T, F] (gradient w.r.t. PReLU(IN(raw))), dslope_f [F], dwf, dbf (one float each)
extern "C" int seb200_mask_tail_bwd(const float* mask_raw, const float* mask_stats, int B, long long rows_per_b, int F, const float* const* scalars5,
                                    const float* slope_f, const float* in3, const float* dest, float* dp1, float* dslope_f, float* dwf, float* dbf,
                                    float* workspace, void* stream) {
  SEB_REQUIRE(mask_raw && mask_stats && scalars5 && slope_f && in3 && dest && dp1 && dslope_f && dwf && dbf && workspace && B > 0 && B < 65536 && rows_per_b > 0 &&
              F > 0 && F <= 254, SEB_EINVAL, "mask_tail_bwd: bad arguments");
  MaskScal ms;
  for (int i = 0; i < 5; ++i) { SEB_REQUIRE(scalars5[i], SEB_EINVAL, "mask_tail_bwd: null scalar %d", i); ms.p[i] = scalars5[i]; }
  int chunks = (int)((rows_per_b + 15) / 16);
  const int cap = 148 * 4 / B + 1;
  if (chunks > cap) chunks = cap;
  const int rows_per_chunk = (int)((rows_per_b + chunks - 1) / chunks);
  chunks = (int)((rows_per_b + rows_per_chunk - 1) / rows_per_chunk);
  dim3 grid(chunks, B);
  mask_tail_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(mask_raw, mask_stats, rows_per_b, F, rows_per_chunk, ms, slope_f, in3, dest, dp1, workspace);
  SEB_CHECK_LAUNCH("mask_tail_bwd_kernel");
  // partial rows are (dslope_f [F] | dwf | dbf): dwf and dbf are adjacent single floats only if the caller made them so; finish in two steps
  finish_f32_kernel<<<fin_grid(F + 2), 256, 0, ST(stream)>>>(workspace, B * chunks, F + 2, dslope_f, F, workspace + (long long)B * chunks * (F + 2));
  SEB_CHECK_LAUNCH("finish_f32_kernel");
  cudaMemcpyAsync(dwf, workspace + (long long)B * chunks * (F + 2), sizeof(float), cudaMemcpyDeviceToDevice, ST(stream));
  cudaMemcpyAsync(dbf, workspace + (long long)B * chunks * (F + 2) + 1, sizeof(float), cudaMemcpyDeviceToDevice, ST(stream));
  return 0;
}

extern "C" int seb200_conv1x1_in3_wgrad(const float* in3, const float* g, long long pixels, float* dw, float* db, float* workspace, void* stream) {
  SEB_REQUIRE(in3 && g && dw && db && workspace && pixels > 0 && aligned16(g), SEB_EINVAL, "conv1x1_in3_wgrad: bad arguments");
  const int nb = tgrid(pixels, 16 * 16, 148 * 4);
  conv1x1_in3_wgrad_kernel<<<nb, 256, 0, ST(stream)>>>(in3, g, pixels, workspace);
  SEB_CHECK_LAUNCH("conv1x1_in3_wgrad_kernel");
  finish_f32_kernel<<<fin_grid(256), 256, 0, ST(stream)>>>(workspace, nb, 256, dw, 192, db);
  SEB_CHECK_LAUNCH("finish_f32_kernel");
  return 0;
}

extern "C" int seb200_merge_ri(const float* re, const float* im, long long n, float* est, void* stream) {
  SEB_REQUIRE(re && im && est && n > 0, SEB_EINVAL, "merge_ri: bad arguments");
  merge_ri_kernel<<<tgrid(n, 256 * 4), 256, 0, ST(stream)>>>(re, im, n, reinterpret_cast<float2*>(est));
  SEB_CHECK_LAUNCH("merge_ri_kernel");
  return 0;
}

extern "C" int seb200_qkv_to_f16(const float* qkv, long long M, void* out, void* stream) {
  SEB_REQUIRE(qkv && out && M > 0 && aligned16(qkv) && aligned16(out), SEB_EINVAL, "qkv_to_f16: bad arguments");
  qkv_to_f16_kernel<<<tgrid(M * 48, 256 * 4), 256, 0, ST(stream)>>>(qkv, M, reinterpret_cast<__half*>(out));
  SEB_CHECK_LAUNCH("qkv_to_f16_kernel");
  return 0;
}
