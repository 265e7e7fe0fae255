// dsp_bwd.cu -- backward of the power-compressed STFT / iSTFT bracket (the consistency-loss chain of train_gan,
// core/function.py:231-254, differentiates through uncompressed_istft -> compressed_stft).  The two DFT contractions of the backward
// pass run on the GEMM engine with the transposed bases (the adjoint of a dense contraction is a dense contraction); this file holds
// the bandwidth-bound pieces around them.
//
//   compressed_stft      x --reflect pad, frame--> frames --DFT basis--> X --|X|^0.3 e^{j arg X}--> Y
//     backward           gY --Jacobian of the compression--> gX rows --basis^T (GEMM)--> gFrames --adjoint of frame + reflect pad--> gx
//   uncompressed_istft   Y --|Y|^(1/0.3) e^{j arg Y}--> Z rows --iDFT basis--> frames --overlap-add / envelope--> y
//     backward           gy --/ envelope, zero pad, frame (Hankel loader)--> gFrames --basis^T (GEMM)--> gZ rows --Jacobian--> gY
//
// Both non-linearities are  w = v |v|^(p-1)  on complex v (p = 0.3 or 1/0.3).  With u = v / |v| (as a real 2-vector) the real
// Jacobian is |v|^(p-1) (I + (p-1) u u^T), symmetric, so  gv = |v|^(p-1) (g + (p-1) (g.u) u).   PyTorch's convention for a complex
// tensor in a real loss is grad = dL/dRe + j dL/dIm, which is exactly the pair this file reads and writes.
#include "common.cuh"

namespace seb {

// gv for the compression: v = X is not stored; from Y = X |X|^(c-1):  |X| = |Y|^(1/c), u = Y / |Y|, |X|^(c-1) = |Y|^((c-1)/c)
__device__ __forceinline__ float2 compress_jacobian(float2 y, float2 g) {
  const float m2 = y.x * y.x + y.y * y.y;
  if (!(m2 > 0.f)) return make_float2(0.f, 0.f);            // d|X|^0.3 at 0 is unbounded; the reference's autograd yields nan there
  const float c = 0.3f;
  const float ry = sqrtf(m2), inv = 1.0f / ry;
  const float2 u = make_float2(y.x * inv, y.y * inv);
  const float s = powf(ry, (c - 1.0f) / c);
  const float gu = (c - 1.0f) * (g.x * u.x + g.y * u.y);
  return make_float2(s * fmaf(gu, u.x, g.x), s * fmaf(gu, u.y, g.y));
}

// gY for the decompression Z = Y |Y|^(e-1), e = 1/0.3
__device__ __forceinline__ float2 decompress_jacobian(float2 y, float2 g) {
  const float m2 = y.x * y.x + y.y * y.y;
  if (!(m2 > 0.f)) return make_float2(0.f, 0.f);
  const float e = 1.0f / 0.3f;
  const float inv = rsqrtf(m2);
  const float2 u = make_float2(y.x * inv, y.y * inv);
  const float s = powf(m2, 0.5f * (e - 1.0f));
  const float gu = (e - 1.0f) * (g.x * u.x + g.y * u.y);
  return make_float2(s * fmaf(gu, u.x, g.x), s * fmaf(gu, u.y, g.y));
}

// (B, F, T) complex Y and gY -> rows [B*T, ldz] = (gX_re, gX_im) per bin, zero padded (the A operand of the basis^T GEMM)
__global__ void __launch_bounds__(256) compress_backward_rows_kernel(const float2* __restrict__ spec, const float2* __restrict__ gspec,
                                                                    int F, int T, float* __restrict__ rows, int ldz) {
  __shared__ float2 ty_[32][33], tg_[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, tyi = threadIdx.x >> 5;
  for (int i = tyi; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    const bool ok = f < F && t < T;
    const long long o = ((long long)b * F + f) * T + t;
    ty_[i][tx] = ok ? spec[o] : make_float2(0.f, 0.f);
    tg_[i][tx] = ok ? gspec[o] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int i = tyi; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (t < T && 2 * f < ldz)
      *reinterpret_cast<float2*>(rows + ((long long)b * T + t) * ldz + 2 * f) =
          (f < F) ? compress_jacobian(ty_[tx][i], tg_[tx][i]) : make_float2(0.f, 0.f);
  }
}

// rows [B*T, ldz] of gZ + (B, F, T) complex Y -> gY (B, F, T)
__global__ void __launch_bounds__(256) decompress_backward_spec_kernel(const float2* __restrict__ spec, const float* __restrict__ rows, int F, int T,
                                                                      int ldz, float2* __restrict__ gspec) {
  __shared__ float2 tg_[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, tyi = threadIdx.x >> 5;
  for (int i = tyi; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    tg_[i][tx] = (t < T && f < F) ? *reinterpret_cast<const float2*>(rows + ((long long)b * T + t) * ldz + 2 * f) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int i = tyi; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    if (f < F && t < T) {
      const long long o = ((long long)b * F + f) * T + t;
      gspec[o] = decompress_jacobian(spec[o], tg_[tx][i]);
    }
  }
}

// adjoint of (reflect pad 200 | frame with hop 100, width 400) as a gather: G(p) = sum_t gframes[b, t, p - 100 t] is the gradient
// of the padded sample p; the interior sample i collects G(200 + i) and the (at most two) padded positions that mirror it.
__global__ void __launch_bounds__(256) stft_fold_kernel(const float* __restrict__ gframes, int T, int ldf, int L, float* __restrict__ gx) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const float* fr = gframes + (long long)b * T * ldf;
  auto G = [&](int p) {
    int t_hi = p / 100; if (t_hi > T - 1) t_hi = T - 1;
    int t_lo = (p - 399 + 99) / 100; if (t_lo < 0) t_lo = 0;
    float acc = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) acc += fr[(long long)t * ldf + (p - 100 * t)];
    return acc;
  };
  float v = G(200 + i);
  if (i >= 1 && i <= 200) v += G(200 - i);                             // left mirror: xpad[p] = x[200 - p], p < 200
  if (i >= L - 201 && i <= L - 2) v += G(200 + 2 * (L - 1) - i);       // right mirror: xpad[200 + L + k] = x[L - 2 - k]
  gx[(long long)b * L + i] = v;
}

// wpad[b, 0 : Lout + 400] = zero-pad200(gy[b, :] * inv_env): framing it with hop 100 / width 400 (the Hankel loader) is the
// adjoint of the overlap-add + envelope division + trim of torch.istft
__global__ void __launch_bounds__(256) istft_grad_pad_kernel(const float* __restrict__ gy, const float* __restrict__ inv_env, int Lout,
                                                            float* __restrict__ wpad) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Lout + 400) return;
  const int m = q - 200;
  wpad[(long long)b * (Lout + 400) + q] = (m >= 0 && m < Lout) ? gy[(long long)b * Lout + m] * inv_env[m] : 0.f;
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_compress_backward_rows(const float* spec_ri, const float* gspec_ri, int B, int F, int T, float* rows, int ldz, void* stream) {
  SEB_REQUIRE(spec_ri && gspec_ri && rows && B > 0 && B < 65536 && F > 0 && T > 0 && ldz >= 2 * F && ldz % 4 == 0, SEB_EINVAL, "compress_backward_rows: bad arguments");
  dim3 grid((T + 31) / 32, (ldz / 2 + 31) / 32, B);
  compress_backward_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec_ri), reinterpret_cast<const float2*>(gspec_ri), F, T, rows, ldz);
  SEB_CHECK_LAUNCH("compress_backward_rows_kernel");
  return 0;
}

extern "C" int seb200_decompress_backward_spec(const float* spec_ri, const float* rows, int B, int F, int T, int ldz, float* gspec_ri, void* stream) {
  SEB_REQUIRE(spec_ri && gspec_ri && rows && B > 0 && B < 65536 && F > 0 && T > 0 && ldz >= 2 * F && ldz % 2 == 0, SEB_EINVAL, "decompress_backward_spec: bad arguments");
  dim3 grid((T + 31) / 32, (F + 31) / 32, B);
  decompress_backward_spec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec_ri), rows, F, T, ldz, reinterpret_cast<float2*>(gspec_ri));
  SEB_CHECK_LAUNCH("decompress_backward_spec_kernel");
  return 0;
}

extern "C" int seb200_stft_fold(const float* gframes, int B, int T, int ldf, int L, float* gx, void* stream) {
  SEB_REQUIRE(gframes && gx && B > 0 && B < 65536 && T > 0 && ldf >= 400 && L > 200 && L == 100 * (T - 1), SEB_EINVAL,
              "stft_fold: bad arguments (L must be 100*(T-1) > 200)");
  dim3 grid((L + 255) / 256, B);
  stft_fold_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gframes, T, ldf, L, gx);
  SEB_CHECK_LAUNCH("stft_fold_kernel");
  return 0;
}

extern "C" int seb200_istft_grad_pad(const float* gy, const float* inv_env, int B, int Lout, float* wpad, void* stream) {
  SEB_REQUIRE(gy && inv_env && wpad && B > 0 && B < 65536 && Lout > 0 && Lout % 100 == 0, SEB_EINVAL, "istft_grad_pad: bad arguments");
  dim3 grid((Lout + 400 + 255) / 256, B);
  istft_grad_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gy, inv_env, Lout, wpad);
  SEB_CHECK_LAUNCH("istft_grad_pad_kernel");
  return 0;
}
