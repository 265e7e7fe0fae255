// merge.cu -- the diffusion-step branch of MergeBlock (models/tsc_diffusion.py:16-41).
//
// MergeBlock.forward is   y = W_m (x + d) + b_m + W_c cond + b_c;  out = (x + W_o (sigmoid(y_g) * tanh(y_f)) + b_o) / sqrt(2)
// with d = diffusion_projection(DiffusionEmbedding(step)) constant over an utterance.  The two token-wise contractions run on
// the persistent tcgen05 token GEMM (tok_gemm.cu: SEB_LOAD_ROWS2 + SEB_EPI_GATE, then SEB_EPI_RESID_SCALE); this file holds the
// only other piece: the step embedding MLP (models/DiffuSE.py:46-62), evaluated once per forward for 1 or B steps, and its
// projection through W_m so that the GEMM sees it as a per-utterance bias row:  W_m (x + d) = W_m x + (W_m d).
#include "common.cuh"

namespace seb {

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// out[j] = act(b[j] + sum_k w[j, k] * in[k]) for j < nout: one warp per output, lanes stride k (coalesced rows of w)
template <bool SILU>
__device__ __forceinline__ void dense_rows(const float* __restrict__ w, const float* __restrict__ b, const float* in, float* out,
                                           int nout, int nin, int warp, int lane, int nwarps) {
  for (int j = warp; j < nout; j += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < nin; k += 32) acc = fmaf(__ldg(w + (long long)j * nin + k), in[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + (b ? __ldg(b + j) : 0.f);
      out[j] = SILU ? silu_exact(v) : v;
    }
  }
}

__global__ void __launch_bounds__(512) diffusion_embed_kernel(const float* __restrict__ steps, const float* __restrict__ table, int max_steps,
                                                              const float* __restrict__ w1, const float* __restrict__ b1,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              const float* __restrict__ wp, const float* __restrict__ bp,
                                                              const float* __restrict__ wm, float* __restrict__ d_out,
                                                              float* __restrict__ rowbias) {
  __shared__ float e[128], h1[512], h2[512], d[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, s = blockIdx.x;
  // DiffuSE.py:46-50,57-62: integer steps index the table; fractional steps interpolate between the neighbouring rows
  // (for an integral value floor == ceil and the interpolation returns the row itself)
  const float t = steps[s];
  int lo = (int)floorf(t), hi = (int)ceilf(t);
  lo = min(max(lo, 0), max_steps - 1);
  hi = min(max(hi, 0), max_steps - 1);
  if (tid < 128) {
    const float a = table[lo * 128 + tid], b = table[hi * 128 + tid];
    e[tid] = a + (b - a) * (t - (float)lo);
  }
  __syncthreads();
  dense_rows<true>(w1, b1, e, h1, 512, 128, warp, lane, 16);
  __syncthreads();
  dense_rows<true>(w2, b2, h1, h2, 512, 512, warp, lane, 16);
  __syncthreads();
  dense_rows<false>(wp, bp, h2, d, 64, 512, warp, lane, 16);
  __syncthreads();
  if (tid < 64) d_out[s * 64 + tid] = d[tid];
  dense_rows<false>(wm, nullptr, d, rowbias + s * 128, 128, 64, warp, lane, 16);
}

// one reverse-diffusion update of the waveform (inference_diffuse.py:255-264):
//   out[b, i] = (ca * audio[b, i] + cb * noisy[b, i] + cc * pred[b, i] + cs * noise[b, i]) * (c_div ? 1 / c_div[b] : 1)
__global__ void __launch_bounds__(256) diffusion_update_kernel(const float* __restrict__ audio, const float* __restrict__ noisy, long long ldn,
                                                               const float* __restrict__ pred, const float* __restrict__ noise, long long L,
                                                               float ca, float cb, float cc, float cs, const float* __restrict__ c_div,
                                                               float* __restrict__ out) {
  const int b = blockIdx.y;
  const float sc = c_div ? 1.0f / c_div[b] : 1.0f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < L; i += (long long)gridDim.x * 256) {
    const long long o = (long long)b * L + i;
    float v = ca * audio[o] + cb * noisy[(long long)b * ldn + i] + cc * pred[o];
    if (noise) v = fmaf(cs, noise[o], v);
    out[o] = v * sc;
  }
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_diffusion_update(const float* audio, const float* noisy, long long ld_noisy, const float* pred, const float* noise,
                                       int B, long long L, float ca, float cb, float cc, float cs, const float* c_div, float* out, void* stream) {
  SEB_REQUIRE(audio && noisy && pred && out && B > 0 && B < 65536 && L > 0 && ld_noisy >= L, SEB_EINVAL, "diffusion_update: bad arguments");
  dim3 grid((unsigned)((L + 255) / 256 < 1024 ? (L + 255) / 256 : 1024), B);
  diffusion_update_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(audio, noisy, ld_noisy, pred, noise, L, ca, cb, cc, cs, c_div, out);
  SEB_CHECK_LAUNCH("diffusion_update_kernel");
  return 0;
}

extern "C" int seb200_diffusion_embed(const float* steps, int nsteps, const float* table, int max_steps,
                                      const float* w1, const float* b1, const float* w2, const float* b2,
                                      const float* wp, const float* bp, const float* wm,
                                      float* d_out, float* rowbias, void* stream) {
  SEB_REQUIRE(steps && table && w1 && b1 && w2 && b2 && wp && bp && wm && d_out && rowbias, SEB_EINVAL, "diffusion_embed: null pointer");
  SEB_REQUIRE(nsteps > 0 && max_steps > 0, SEB_EINVAL, "diffusion_embed: nsteps=%d max_steps=%d", nsteps, max_steps);
  diffusion_embed_kernel<<<nsteps, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(steps, table, max_steps, w1, b1, w2, b2, wp, bp, wm, d_out, rowbias);
  SEB_CHECK_LAUNCH("diffusion_embed_kernel");
  return 0;
}
