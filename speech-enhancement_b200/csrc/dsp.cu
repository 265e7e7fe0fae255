// dsp.cu -- the bandwidth-bound pieces of the STFT / iSTFT bracket (the DFTs themselves run on the GEMM engine).
#include "common.cuh"

namespace seb {

// predict() glue (inference_gan.py:79-87) + torch.stft's centre reflect padding, one CTA per utterance.
// With c_in the per-utterance gain is taken from the caller instead (normalize_batch applies the NOISY signal's gain to
// the clean signal as well, core/function.py:647-659).
__global__ void __launch_bounds__(512) rms_pad_kernel(const float* __restrict__ wave, int L, int Lp, int normalize,
                                                     const float* __restrict__ c_in, float* __restrict__ xpad, float* __restrict__ c_out) {
  const int b = blockIdx.x;
  const float* x = wave + (long long)b * L;
  __shared__ double red[16];
  __shared__ float c_s;
  float c = c_in ? c_in[b] : 1.0f;
  if (normalize && !c_in) {
    double acc = 0.0;
    for (int i = threadIdx.x * 4; i < L; i += blockDim.x * 4) {   // short fp32 runs, fp64 across runs
      float s = 0.f;
      for (int j = 0; j < 4 && i + j < L; ++j) s = fmaf(x[i + j], x[i + j], s);
      acc += (double)s;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      c_s = sqrtf((float)L / (float)t);
    }
    __syncthreads();
    c = c_s;
  }
  if (threadIdx.x == 0 && c_out) c_out[b] = c;
  const int total = Lp + 400;
  float* o = xpad + (long long)b * total;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    int j = i - 200;
    if (j < 0) j = -j;                       // reflect, edge sample not repeated
    if (j >= Lp) j = 2 * (Lp - 1) - j;
    if (j >= L) j -= L;                      // tail wrap-padding with the head of the signal
    o[i] = x[j] * c;
  }
}

// complex64 (B, F, T) <-> in3 [B, T, F, 3]; 32x32 tiles through shared memory so both sides stay coalesced.
__global__ void __launch_bounds__(256) spec_to_in3_kernel(const float2* __restrict__ spec, int F, int T, float* __restrict__ in3) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    if (f < F && t < T) tile[i][tx] = spec[((long long)b * F + f) * T + t];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (f < F && t < T) {
      const float2 v = tile[tx][i];
      float* o = in3 + (((long long)b * T + t) * F + f) * 3;
      o[0] = hypotf(v.x, v.y); o[1] = v.x; o[2] = v.y;
    }
  }
}

__global__ void __launch_bounds__(256) in3_to_spec_kernel(const float* __restrict__ in3, int F, int T, float2* __restrict__ spec) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (f < F && t < T) {
      const float* p = in3 + (((long long)b * T + t) * F + f) * 3;
      tile[i][tx] = make_float2(p[1], p[2]);
    }
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    if (f < F && t < T) spec[((long long)b * F + f) * T + t] = tile[tx][i];
  }
}

// power_uncompress (core/function.py:636-645): z = e * |e|^(1/0.3 - 1)
__device__ __forceinline__ float2 decompress(float re, float im) {
  const float m2 = re * re + im * im;
  if (!(m2 > 0.f)) return make_float2(0.f, 0.f);
  const float s = powf(m2, 0.5f * (1.0f / 0.3f - 1.0f));
  return make_float2(re * s, im * s);
}

__global__ void __launch_bounds__(256) decompress_rows_kernel(const float2* __restrict__ est, long long rows, int F,
                                                             float* __restrict__ z, int ldz) {
  const long long row = blockIdx.x;
  const float2* e = est + row * F;
  float* o = z + row * ldz;
  for (int i = threadIdx.x; i < ldz / 2; i += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (i < F) v = decompress(e[i].x, e[i].y);
    *reinterpret_cast<float2*>(o + 2 * i) = v;
  }
}

__global__ void __launch_bounds__(256) spec_decompress_rows_kernel(const float2* __restrict__ spec, int F, int T,
                                                                  float* __restrict__ z, int ldz) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    tile[i][tx] = (f < F && t < T) ? spec[((long long)b * F + f) * T + t] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (t < T && 2 * f < ldz) {
      const float2 v = tile[tx][i];
      *reinterpret_cast<float2*>(z + ((long long)b * T + t) * ldz + 2 * f) = (f < F) ? decompress(v.x, v.y) : make_float2(0.f, 0.f);
    }
  }
}

// torch.istft's overlap-add as a gather: every output sample sums its <= 4 frames in fixed order.
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ frames, int T, int ldf,
                                                         const float* __restrict__ inv_env, const float* __restrict__ c,
                                                         float* __restrict__ out, int Lout, int ld_out) {
  const int b = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Lout) return;
  const int pos = m + 200;                       // position in the untrimmed signal
  int t_hi = pos / 100; if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = (pos - 399 + 99) / 100; if (t_lo < 0) t_lo = 0;
  float acc = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) acc += frames[((long long)b * T + t) * ldf + (pos - 100 * t)];
  float v = acc * inv_env[m];
  if (c) v = v / c[b];
  out[(long long)b * ld_out + m] = v;
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_rms_pad(const float* wave, int B, int L, int Lp, int normalize, float* xpad, float* c_out, void* stream) {
  SEB_REQUIRE(wave && xpad && c_out && B > 0 && L > 0 && Lp >= L && Lp % 100 == 0 && Lp - L < 100 && Lp > 200, SEB_EINVAL,
              "rms_pad: bad arguments B=%d L=%d Lp=%d", B, L, Lp);
  rms_pad_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(wave, L, Lp, normalize, nullptr, xpad, c_out);
  SEB_CHECK_LAUNCH("rms_pad_kernel");
  return 0;
}

extern "C" int seb200_scale_pad(const float* wave, int B, int L, int Lp, const float* c_in, float* xpad, void* stream) {
  SEB_REQUIRE(wave && xpad && c_in && B > 0 && L > 0 && Lp >= L && Lp % 100 == 0 && Lp - L < 100 && Lp > 200, SEB_EINVAL,
              "scale_pad: bad arguments B=%d L=%d Lp=%d", B, L, Lp);
  rms_pad_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(wave, L, Lp, 0, c_in, xpad, nullptr);
  SEB_CHECK_LAUNCH("rms_pad_kernel");
  return 0;
}

extern "C" int seb200_spec_to_in3(const float* spec_ri, int B, int F, int T, float* in3, void* stream) {
  SEB_REQUIRE(spec_ri && in3 && B > 0 && F > 0 && T > 0 && B < 65536, SEB_EINVAL, "spec_to_in3: bad arguments");
  dim3 grid((T + 31) / 32, (F + 31) / 32, B);
  spec_to_in3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec_ri), F, T, in3);
  SEB_CHECK_LAUNCH("spec_to_in3_kernel");
  return 0;
}

extern "C" int seb200_in3_to_spec(const float* in3, int B, int F, int T, float* spec_ri, void* stream) {
  SEB_REQUIRE(spec_ri && in3 && B > 0 && F > 0 && T > 0 && B < 65536, SEB_EINVAL, "in3_to_spec: bad arguments");
  dim3 grid((T + 31) / 32, (F + 31) / 32, B);
  in3_to_spec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in3, F, T, reinterpret_cast<float2*>(spec_ri));
  SEB_CHECK_LAUNCH("in3_to_spec_kernel");
  return 0;
}

extern "C" int seb200_decompress_rows(const float* est, int rows, int F, float* z, int ldz, void* stream) {
  SEB_REQUIRE(est && z && rows > 0 && F > 0 && ldz >= 2 * F && ldz % 4 == 0, SEB_EINVAL, "decompress_rows: bad arguments");
  decompress_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(est), rows, F, z, ldz);
  SEB_CHECK_LAUNCH("decompress_rows_kernel");
  return 0;
}

extern "C" int seb200_spec_decompress_rows(const float* spec_ri, int B, int F, int T, float* z, int ldz, void* stream) {
  SEB_REQUIRE(spec_ri && z && B > 0 && B < 65536 && F > 0 && T > 0 && ldz >= 2 * F && ldz % 4 == 0, SEB_EINVAL, "spec_decompress_rows: bad arguments");
  dim3 grid((T + 31) / 32, (ldz / 2 + 31) / 32, B);
  spec_decompress_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec_ri), F, T, z, ldz);
  SEB_CHECK_LAUNCH("spec_decompress_rows_kernel");
  return 0;
}

extern "C" int seb200_overlap_add(const float* frames, int B, int T, int ldf, const float* inv_env, const float* c,
                                  float* out, int Lout, int ld_out, void* stream) {
  SEB_REQUIRE(frames && inv_env && out && B > 0 && B < 65536 && T > 0 && ldf >= 400 && Lout == 100 * (T - 1) && ld_out >= Lout, SEB_EINVAL,
              "overlap_add: bad arguments (Lout must be 100*(T-1))");
  dim3 grid((Lout + 255) / 256, B);
  overlap_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, T, ldf, inv_env, c, out, Lout, ld_out);
  SEB_CHECK_LAUNCH("overlap_add_kernel");
  return 0;
}
