// norm_act.cu -- HBM-bound kernels of the encoder / decoders / conformer glue: pointwise conv on the 3-channel
// input, InstanceNorm statistics + apply (+PReLU), the 1- and 2-channel output heads, mask recombination and the
// post-norm + residual.  All are single-pass, float4-vectorised, channels-last.
#include "common.cuh"

namespace seb {

// ---- DenseEncoder.conv_1[0]: 1x1 conv 3 -> 64 (generator.py:39).  16 threads per pixel, 4 channels each.
__global__ void __launch_bounds__(256) conv1x1_in3_kernel(const float* __restrict__ in3, long long pixels,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out) {
  const int cq = threadIdx.x & 15;
  float wr[4][3], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cq * 4 + j;
    wr[j][0] = w[c * 3 + 0]; wr[j][1] = w[c * 3 + 1]; wr[j][2] = w[c * 3 + 2];
    br[j] = bias ? bias[c] : 0.f;
  }
  for (long long p = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); p < pixels; p += (long long)gridDim.x * 16) {
    const float a = in3[p * 3], b = in3[p * 3 + 1], c = in3[p * 3 + 2];
    float4 o;
    o.x = fmaf(wr[0][2], c, fmaf(wr[0][1], b, fmaf(wr[0][0], a, br[0])));
    o.y = fmaf(wr[1][2], c, fmaf(wr[1][1], b, fmaf(wr[1][0], a, br[1])));
    o.z = fmaf(wr[2][2], c, fmaf(wr[2][1], b, fmaf(wr[2][0], a, br[2])));
    o.w = fmaf(wr[3][2], c, fmaf(wr[3][1], b, fmaf(wr[3][0], a, br[3])));
    st4(out + p * 64 + cq * 4, o);
  }
}

// ---- InstanceNorm statistics, deterministic two-stage reduction -------------------------------
// stage 1: grid (chunks, B); each CTA reduces `rows_per_chunk` pixels of one utterance for all C channels.
// fp32 inside a thread's short run, fp64 across runs / threads / chunks (no atomics => bitwise reproducible).
constexpr int IN_ROWS_PER_CHUNK = 1024;

template <int C>   // C = 64: 16 float4 lanes per pixel, 16 pixel lanes per CTA
__global__ void __launch_bounds__(256) inorm_partial_kernel(const float* __restrict__ x, long long pix_per_b,
                                                           double* __restrict__ part) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int cq = threadIdx.x & 15, pl = threadIdx.x >> 4;
  const long long p0 = (long long)chunk * IN_ROWS_PER_CHUNK;
  long long p1 = p0 + IN_ROWS_PER_CHUNK; if (p1 > pix_per_b) p1 = pix_per_b;
  const float* xb = x + (long long)b * pix_per_b * C;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (long long p = p0 + pl; p < p1; p += 16 * 8) {
    float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long pp = p + u * 16;
      if (pp < p1) {
        const float4 v = ldg4(xb + pp * C + cq * 4);
        fs[0] += v.x; fs[1] += v.y; fs[2] += v.z; fs[3] += v.w;
        fq[0] = fmaf(v.x, v.x, fq[0]); fq[1] = fmaf(v.y, v.y, fq[1]); fq[2] = fmaf(v.z, v.z, fq[2]); fq[3] = fmaf(v.w, v.w, fq[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[j] += (double)fs[j]; q[j] += (double)fq[j]; }
  }
  __shared__ double sh[16][64][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[pl][cq * 4 + j][0] = s[j]; sh[pl][cq * 4 + j][1] = q[j]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x >> 1, k = threadIdx.x & 1;
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += sh[i][c][k];
    part[(((long long)b * gridDim.x + chunk) * C + c) * 2 + k] = t;
  }
}

// C = 1 (MaskDecoder.norm): plain strided reduction over the chunk
__global__ void __launch_bounds__(256) inorm_partial1_kernel(const float* __restrict__ x, long long pix_per_b, double* __restrict__ part) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const long long p0 = (long long)chunk * IN_ROWS_PER_CHUNK * 64;
  long long p1 = p0 + (long long)IN_ROWS_PER_CHUNK * 64; if (p1 > pix_per_b) p1 = pix_per_b;
  const float* xb = x + (long long)b * pix_per_b;
  double s = 0, q = 0;
  for (long long p = p0 + threadIdx.x; p < p1; p += 256 * 8) {
    float fs = 0, fq = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { const long long pp = p + u * 256; if (pp < p1) { const float v = xb[pp]; fs += v; fq = fmaf(v, v, fq); } }
    s += fs; q += fq;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ double sh[8][2];
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = s; sh[threadIdx.x >> 5][1] = q; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    part[((long long)b * gridDim.x + chunk) * 2 + threadIdx.x] = t;
  }
}

__global__ void inorm_finalize_kernel(const double* __restrict__ part, int chunks, int C, long long pix_per_b, float* __restrict__ stats) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int k = 0; k < chunks; ++k) { const double* p = part + (((long long)b * chunks + k) * C + c) * 2; s += p[0]; q += p[1]; }
  const double mean = s / (double)pix_per_b;
  double var = q / (double)pix_per_b - mean * mean;
  if (var < 0) var = 0;
  stats[((long long)b * C + c) * 2] = (float)mean;
  stats[((long long)b * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
}

// ---- y = PReLU(gamma * (x - mean) * rstd + beta), C = 64 ------------------------------------
// SPLIT = false: fp32 output.  SPLIT = true: the pre-split conv-input format, per pixel 64 bf16 hi | 64 bf16 lo.
__device__ __forceinline__ void store_split4(uint8_t* pixel_base, int cq, float4 v) {
  uint32_t h0, l0, h1, l1;
  split_bf16x2(v.x, v.y, h0, l0);
  split_bf16x2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(pixel_base + cq * 8) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(pixel_base + 128 + cq * 8) = make_uint2(l0, l1);
}

template <bool SPLIT>
__global__ void __launch_bounds__(256) inorm_prelu_kernel(const float* __restrict__ x, long long pix_per_b,
                                                         const float* __restrict__ stats, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ slope,
                                                         void* __restrict__ yv) {
  const int b = blockIdx.y, cq = threadIdx.x & 15;
  float sc[4], sh[4], sl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cq * 4 + j;
    const float mean = stats[((long long)b * 64 + c) * 2], rstd = stats[((long long)b * 64 + c) * 2 + 1];
    sc[j] = rstd * gamma[c];
    sh[j] = beta[c] - mean * sc[j];
    sl[j] = slope[c];
  }
  const float* xb = x + (long long)b * pix_per_b * 64;
  for (long long p = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); p < pix_per_b; p += (long long)gridDim.x * 16) {
    float4 v = ldg4(xb + p * 64 + cq * 4);
    v.x = fmaf(v.x, sc[0], sh[0]); v.y = fmaf(v.y, sc[1], sh[1]); v.z = fmaf(v.z, sc[2], sh[2]); v.w = fmaf(v.w, sc[3], sh[3]);
    v.x = v.x >= 0.f ? v.x : v.x * sl[0]; v.y = v.y >= 0.f ? v.y : v.y * sl[1];
    v.z = v.z >= 0.f ? v.z : v.z * sl[2]; v.w = v.w >= 0.f ? v.w : v.w * sl[3];
    const long long gp = (long long)b * pix_per_b + p;
    if (SPLIT) store_split4(reinterpret_cast<uint8_t*>(yv) + gp * 256, cq, v);
    else st4(reinterpret_cast<float*>(yv) + gp * 64 + cq * 4, v);
  }
}

__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long long pixels, uint8_t* __restrict__ y) {
  const int cq = threadIdx.x & 15;
  for (long long p = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); p < pixels; p += (long long)gridDim.x * 16)
    store_split4(y + p * 256, cq, ldg4(x + p * 64 + cq * 4));
}

// ---- MaskDecoder.conv_1: Conv2d(64 -> 1, (1,2)); one warp per output pixel pair of rows ------------
// x [rows, Fin, 64]; out [rows, Fin-1].  16 lanes per output pixel (float4 each, both taps), shuffle reduce.
__global__ void __launch_bounds__(256) mask_conv_kernel(const float* __restrict__ x, long long rows, int Fin,
                                                       const float* __restrict__ w, float bias, float* __restrict__ out) {
  const int cq = threadIdx.x & 15;
  const float4 w0 = ldg4(w + cq * 4), w1 = ldg4(w + 64 + cq * 4);
  const int Fo = Fin - 1;
  const long long total = rows * Fo;
  for (long long ob_ = (long long)blockIdx.x * 16; ob_ < total; ob_ += (long long)gridDim.x * 16) {   // block-uniform trip count (shuffles inside)
    const long long o = ob_ + (threadIdx.x >> 4);
    float acc = 0.f;
    const bool ok = o < total;
    if (ok) {
      const long long r = o / Fo; const int f = (int)(o - r * Fo);
      const float* p = x + (r * Fin + f) * 64 + cq * 4;
      const float4 a = ldg4(p), b = ldg4(p + 64);
      acc = a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w + b.x * w1.x + b.y * w1.y + b.z * w1.z + b.w * w1.w;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4); acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    if (ok && cq == 0) out[o] = acc + bias;
  }
}

// ---- ComplexDecoder tail: IN(64)+PReLU(64) on load, Conv2d(64 -> 2, (1,2)) (generator.py:127-128) ----
__global__ void __launch_bounds__(256) complex_conv_kernel(const float* __restrict__ x, long long rows_per_b, int Fin,
                                                          const float* __restrict__ stats, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ slope,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ out) {
  const int b = blockIdx.y, cq = threadIdx.x & 15;
  float sc[4], sh[4], sl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cq * 4 + j;
    const float mean = stats[((long long)b * 64 + c) * 2], rstd = stats[((long long)b * 64 + c) * 2 + 1];
    sc[j] = rstd * gamma[c]; sh[j] = beta[c] - mean * sc[j]; sl[j] = slope[c];
  }
  // w: [out 2][tap 2][64]
  const float4 w00 = ldg4(w + cq * 4), w01 = ldg4(w + 64 + cq * 4), w10 = ldg4(w + 128 + cq * 4), w11 = ldg4(w + 192 + cq * 4);
  const float b0 = bias[0], b1 = bias[1];
  const int Fo = Fin - 1;
  const long long total = rows_per_b * Fo;
  const float* xb = x + (long long)b * rows_per_b * Fin * 64;
  float* ob = out + (long long)b * total * 2;
  auto act = [&](float4 v) {
    v.x = fmaf(v.x, sc[0], sh[0]); v.y = fmaf(v.y, sc[1], sh[1]); v.z = fmaf(v.z, sc[2], sh[2]); v.w = fmaf(v.w, sc[3], sh[3]);
    v.x = v.x >= 0.f ? v.x : v.x * sl[0]; v.y = v.y >= 0.f ? v.y : v.y * sl[1];
    v.z = v.z >= 0.f ? v.z : v.z * sl[2]; v.w = v.w >= 0.f ? v.w : v.w * sl[3];
    return v;
  };
  for (long long ob_ = (long long)blockIdx.x * 16; ob_ < total; ob_ += (long long)gridDim.x * 16) {   // block-uniform trip count (shuffles inside)
    const long long o = ob_ + (threadIdx.x >> 4);
    float a0 = 0.f, a1 = 0.f;
    const bool ok = o < total;
    if (ok) {
      const long long r = o / Fo; const int f = (int)(o - r * Fo);
      const float* p = xb + (r * Fin + f) * 64 + cq * 4;
      const float4 u = act(ldg4(p)), v = act(ldg4(p + 64));
      a0 = u.x * w00.x + u.y * w00.y + u.z * w00.z + u.w * w00.w + v.x * w01.x + v.y * w01.y + v.z * w01.z + v.w * w01.w;
      a1 = u.x * w10.x + u.y * w10.y + u.z * w10.z + u.w * w10.w + v.x * w11.x + v.y * w11.y + v.z * w11.z + v.w * w11.w;
    }
#pragma unroll
    for (int s = 1; s < 16; s <<= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, s); a1 += __shfl_xor_sync(0xffffffffu, a1, s); }
    if (ok && cq == 0) *reinterpret_cast<float2*>(ob + o * 2) = make_float2(a0 + b0, a1 + b1);
  }
}

// ---- mask tail + recombination (generator.py:110-112, 158-165) ----
__global__ void __launch_bounds__(256) mask_recombine_kernel(const float* __restrict__ raw, const float* __restrict__ stats,
                                                            long long rows_per_b, int F, float in_gamma, float in_beta, float slope1,
                                                            float wf, float bf, const float* __restrict__ slope_f,
                                                            const float* __restrict__ in3, const float* __restrict__ cplx,
                                                            float* __restrict__ est, float* __restrict__ mask_out) {
  const int b = blockIdx.y;
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
    if (mask_out) mask_out[base + i] = v;
  }
}

__global__ void __launch_bounds__(256) split_ri_kernel(const float2* __restrict__ est, long long n, float* __restrict__ re, float* __restrict__ im) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = est[i];
    re[i] = v.x; im[i] = v.y;
  }
}

// ---- out = LayerNorm(x) * g + b + resid  (post_norm + TSCB outer residual); 16 lanes per 64-wide token ----
__global__ void __launch_bounds__(256) layernorm_residual_kernel(const float* __restrict__ x, long long tokens,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ resid, float* __restrict__ out) {
  const int cq = threadIdx.x & 15;
  const float4 g = ldg4(gamma + cq * 4), bb = ldg4(beta + cq * 4);
  for (long long tb_ = (long long)blockIdx.x * 16; tb_ < tokens; tb_ += (long long)gridDim.x * 16) {   // block-uniform trip count
    const long long t = tb_ + (threadIdx.x >> 4);
    const bool ok = t < tokens;
    float4 v = ok ? ldg4(x + t * 64 + cq * 4) : make_float4(0, 0, 0, 0);
    float s = v.x + v.y + v.z + v.w;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 64.0f);
    v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
    float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / 64.0f) + 1e-5f);
    if (ok) {
      float4 r = resid ? *reinterpret_cast<const float4*>(resid + t * 64 + cq * 4) : make_float4(0, 0, 0, 0);
      float4 o;
      o.x = v.x * rstd * g.x + bb.x + r.x; o.y = v.y * rstd * g.y + bb.y + r.y;
      o.z = v.z * rstd * g.z + bb.z + r.z; o.w = v.w * rstd * g.w + bb.w + r.w;
      st4(out + t * 64 + cq * 4, o);
    }
  }
}

static int grid_for(long long items, int per_block, int cap = 148 * 16) {
  long long g = (items + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_conv1x1_in3(const float* in3, long long pixels, const float* w, const float* bias, float* out, void* stream) {
  SEB_REQUIRE(in3 && w && out && pixels > 0 && aligned16(out), SEB_EINVAL, "conv1x1_in3: bad arguments");
  conv1x1_in3_kernel<<<grid_for(pixels, 16 * 8), 256, 0, (cudaStream_t)stream>>>(in3, pixels, w, bias, out);
  SEB_CHECK_LAUNCH("conv1x1_in3_kernel");
  return 0;
}

static int in_chunks(long long pix_per_b, int C) {
  const long long per = (C == 1) ? (long long)IN_ROWS_PER_CHUNK * 64 : IN_ROWS_PER_CHUNK;
  return (int)((pix_per_b + per - 1) / per);
}

extern "C" long long seb200_inorm_workspace_bytes(int B, long long pix_per_b, int C) {
  return (long long)B * in_chunks(pix_per_b, C) * C * 2 * (long long)sizeof(double);
}

extern "C" int seb200_inorm_stats(const float* x, int B, long long pix_per_b, int C, float* stats, void* workspace,
                                  long long workspace_bytes, void* stream) {
  SEB_REQUIRE(x && stats && workspace && B > 0 && B < 65536 && pix_per_b > 0 && (C == 64 || C == 1), SEB_EINVAL, "inorm_stats: bad arguments (C must be 64 or 1)");
  SEB_REQUIRE(workspace_bytes >= seb200_inorm_workspace_bytes(B, pix_per_b, C) && aligned16(workspace), SEB_EINVAL, "inorm_stats: workspace too small");
  const int chunks = in_chunks(pix_per_b, C);
  dim3 grid(chunks, B);
  if (C == 64) inorm_partial_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pix_per_b, (double*)workspace);
  else inorm_partial1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, pix_per_b, (double*)workspace);
  SEB_CHECK_LAUNCH("inorm_partial_kernel");
  inorm_finalize_kernel<<<B, 64, 0, (cudaStream_t)stream>>>((const double*)workspace, chunks, C, pix_per_b, stats);
  SEB_CHECK_LAUNCH("inorm_finalize_kernel");
  return 0;
}

extern "C" int seb200_inorm_prelu(const float* x, int B, long long pix_per_b, int C, const float* stats, const float* gamma,
                                  const float* beta, const float* slope, void* y, int out_format, void* stream) {
  SEB_REQUIRE(x && y && stats && gamma && beta && slope && B > 0 && B < 65536 && C == 64 && aligned16(x) && aligned16(y), SEB_EINVAL, "inorm_prelu: bad arguments");
  SEB_REQUIRE(out_format == 0 || out_format == 1, SEB_EINVAL, "inorm_prelu: out_format must be 0 (fp32) or 1 (split bf16)");
  dim3 grid(grid_for(pix_per_b, 16 * 8, 148 * 16 / (B < 16 ? B : 16) + 1), B);
  if (out_format == 1) inorm_prelu_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pix_per_b, stats, gamma, beta, slope, y);
  else inorm_prelu_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pix_per_b, stats, gamma, beta, slope, y);
  SEB_CHECK_LAUNCH("inorm_prelu_kernel");
  return 0;
}

extern "C" int seb200_split_planes(const float* x, long long pixels, void* y, void* stream) {
  SEB_REQUIRE(x && y && pixels > 0 && aligned16(x) && aligned16(y), SEB_EINVAL, "split_planes: bad arguments");
  split_planes_kernel<<<grid_for(pixels, 16 * 8), 256, 0, (cudaStream_t)stream>>>(x, pixels, reinterpret_cast<uint8_t*>(y));
  SEB_CHECK_LAUNCH("split_planes_kernel");
  return 0;
}

extern "C" int seb200_mask_conv(const float* x, long long rows, int Fin, const float* w, float bias, float* out, void* stream) {
  SEB_REQUIRE(x && w && out && rows > 0 && Fin > 1 && aligned16(x) && aligned16(w), SEB_EINVAL, "mask_conv: bad arguments");
  mask_conv_kernel<<<grid_for(rows * (Fin - 1), 16 * 8), 256, 0, (cudaStream_t)stream>>>(x, rows, Fin, w, bias, out);
  SEB_CHECK_LAUNCH("mask_conv_kernel");
  return 0;
}

extern "C" int seb200_complex_conv(const float* x, int B, long long rows_per_b, int Fin, const float* stats, const float* gamma,
                                   const float* beta, const float* slope, const float* w, const float* bias, float* out, void* stream) {
  SEB_REQUIRE(x && stats && gamma && beta && slope && w && bias && out && B > 0 && B < 65536 && rows_per_b > 0 && Fin > 1 && aligned16(x) && aligned16(w), SEB_EINVAL, "complex_conv: bad arguments");
  dim3 grid(grid_for(rows_per_b * (Fin - 1), 16 * 8, 148 * 16 / (B < 16 ? B : 16) + 1), B);
  complex_conv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows_per_b, Fin, stats, gamma, beta, slope, w, bias, out);
  SEB_CHECK_LAUNCH("complex_conv_kernel");
  return 0;
}

extern "C" int seb200_mask_recombine(const float* mask_raw, const float* mask_stats, int B, long long rows_per_b, int F, float in_gamma,
                                     float in_beta, float slope1, float wf, float bf, const float* slope_f, const float* in3,
                                     const float* cplx, float* est, float* mask_out, void* stream) {
  SEB_REQUIRE(mask_raw && mask_stats && slope_f && in3 && cplx && est && B > 0 && B < 65536 && rows_per_b > 0 && F > 0, SEB_EINVAL, "mask_recombine: bad arguments");
  dim3 grid(grid_for(rows_per_b * F, 256 * 4, 148 * 16 / (B < 16 ? B : 16) + 1), B);
  mask_recombine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask_raw, mask_stats, rows_per_b, F, in_gamma, in_beta, slope1, wf, bf,
                                                                 slope_f, in3, cplx, est, mask_out);
  SEB_CHECK_LAUNCH("mask_recombine_kernel");
  return 0;
}

extern "C" int seb200_split_ri(const float* est, long long n, float* re, float* im, void* stream) {
  SEB_REQUIRE(est && re && im && n > 0, SEB_EINVAL, "split_ri: bad arguments");
  split_ri_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(est), n, re, im);
  SEB_CHECK_LAUNCH("split_ri_kernel");
  return 0;
}

extern "C" int seb200_layernorm_residual(const float* x, long long tokens, const float* gamma, const float* beta, const float* resid,
                                         float* out, void* stream) {
  SEB_REQUIRE(x && gamma && beta && out && tokens > 0 && aligned16(x) && aligned16(out), SEB_EINVAL, "layernorm_residual: bad arguments");
  layernorm_residual_kernel<<<grid_for(tokens, 16 * 8), 256, 0, (cudaStream_t)stream>>>(x, tokens, gamma, beta, resid, out);
  SEB_CHECK_LAUNCH("layernorm_residual_kernel");
  return 0;
}
