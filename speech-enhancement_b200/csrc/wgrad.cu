// wgrad.cu -- weight gradients of the generator's dense contractions (training step; loss.backward() at
// /root/reference/core/function.py:274 produces them through ATen's conv / addmm backward kernels).
//
//   dW[n, k] = sum_m G[m, n] * A[m, k]          db[n] = sum_m G[m, n]
//
// G [M, N] is the gradient of the layer's output rows (fp32), A [M, K] is the layer's INPUT exactly as the forward GEMM engine read it
// -- the same Loader<> structs of gemm_engine.cuh (plain rows, rows with LayerNorm fused, implicit-GEMM conv gather over the dense
// block's slot list), so a forward layer and its weight gradient share one description of the operand (SebGemm).
//
// The contraction runs over the M pixels / tokens (10^5 .. 10^7), the output is tiny (N x K <= 256 x 1536): split-M.  CTA (s, kc, nt)
// reduces rows [s * rows_per_split, ...) of the 64 x 64 tile (n-tile nt, K chunk kc) in fp32 registers (16 x 16 threads, 4 x 4 outputs
// each; operands staged 32 rows at a time in shared memory) and writes partial[s][n][k]; seb200_wgrad_finish sums the S partials in a
// fixed order (deterministic, no atomics) and scatters them through a two-level index map into the parameter's own layout (conv weights
// are [Cout, Cin, kt, kf] while K runs (tap, slot, channel)) -- typically straight into the flat gradient buffer the all-reduce sends.
#include "gemm_engine.cuh"

namespace seb {

constexpr int WG_ROWS = 32;          // rows staged per step
constexpr int WG_LD = 68;            // padded row length (floats): conflict-free float4 reads for both operands

template <int LK>
__global__ void __launch_bounds__(256) wgrad_kernel(const GemmArgs g, const float* __restrict__ G, long long ldg, int N, int rows_per_split,
                                                    float* __restrict__ partial, float* __restrict__ partial_b) {
  __shared__ __align__(16) float Gs[WG_ROWS][WG_LD];
  __shared__ __align__(16) float As[WG_ROWS][WG_LD];
  const int tid = threadIdx.x, sub = tid & 7, rloc = tid >> 3;      // staging: 8 lanes per row, 8 floats each
  const int tn = tid >> 4, tk = tid & 15;                            // compute: outputs (n0 + 4 tn .. + 3, k0 + 4 tk .. + 3)
  const int split = blockIdx.x, kc = blockIdx.y, n0 = blockIdx.z * 64;
  const int m_lo = split * rows_per_split;
  const int m_hi = min(g.M, m_lo + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool want_bias = partial_b != nullptr && kc == 0 && tk == 0;

  for (int m0 = m_lo; m0 < m_hi; m0 += WG_ROWS) {
    const int m = m0 + rloc;
    float v[8];
    typename Loader<LK>::Row row;
    Loader<LK>::init_row(g, m < m_hi ? m : g.M, row);              // rows past the split read as zeros
    Loader<LK>::load(g, row, kc, sub, v);
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
    if (m < m_hi && n0 + sub * 8 < N) {
      const float* gp = G + (long long)m * ldg + n0 + sub * 8;
      g0 = ldg4(gp); g1 = ldg4(gp + 4);
    }
    __syncthreads();                                                 // previous step's reads are done
    *reinterpret_cast<float4*>(&As[rloc][sub * 8]) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(&As[rloc][sub * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    *reinterpret_cast<float4*>(&Gs[rloc][sub * 8]) = g0;
    *reinterpret_cast<float4*>(&Gs[rloc][sub * 8 + 4]) = g1;
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < WG_ROWS; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&Gs[r][tn * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&As[r][tk * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
      if (want_bias) { bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w; }
    }
  }
  const int K = g.K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn * 4 + i;
    if (n < N) st4(partial + ((long long)split * N + n) * K + kc * 64 + tk * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  }
  if (want_bias) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + tn * 4 + i < N) partial_b[(long long)split * N + n0 + tn * 4 + i] = bsum[i];
  }
}

// dW[n, k] = sum_s partial[s][n][k] -> dw[n * sn + (k / n1) * s0 + (k % n1) * s1] for k < k_logical;  db[n] = sum_s partial_b[s][n]
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b, int S, int N, int K,
                                                           int k_logical, int n1, long long sn, long long s0, long long s1,
                                                           float* __restrict__ dw, float* __restrict__ db) {
  const long long total = (long long)N * K;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx < total) {
    const int n = (int)(idx / K), k = (int)(idx - (long long)n * K);
    if (k < k_logical) {
      float s = 0.f;
      for (int p = 0; p < S; ++p) s += partial[(long long)p * total + idx];
      dw[(long long)n * sn + (long long)(k / n1) * s0 + (long long)(k % n1) * s1] = s;
    }
  } else if (db != nullptr && idx < total + N) {
    const int n = (int)(idx - total);
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += partial_b[(long long)p * N + n];
    db[n] = s;
  }
}

static GemmArgs wg_args(const SebGemm* s) {
  GemmArgs g;
  for (int i = 0; i < 4; ++i) g.a[i] = s->a[i];
  g.lda = s->lda; g.ln_g = s->ln_gamma; g.ln_b = s->ln_beta;
  g.M = s->M; g.N = s->N; g.K = s->K;
  g.B = s->B; g.T = s->T; g.Fin = s->Fin; g.Fout = s->Fout;
  g.taps_t = s->taps_t; g.dil = s->dil; g.stride_f = s->stride_f; g.nslots = s->nslots;
  g.bias = nullptr; g.out = nullptr; g.ldo = 0; g.resid = nullptr; g.ldr = 0; g.alpha = 1.f;
  return g;
}

}  // namespace seb

using namespace seb;

// Number of row splits seb200_wgrad uses for (M, N, K): enough CTAs for two waves of 148 SMs, at least 256 rows per split.
extern "C" int seb200_wgrad_splits(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int tiles = (K / 64) * ((N + 63) / 64);
  int S = (2 * 148 + tiles - 1) / tiles;
  const int maxS = (M + 255) / 256;
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  return S;
}

// Workspace (floats) of one seb200_wgrad call: S * (N * K + N).
extern "C" long long seb200_wgrad_workspace_floats(int M, int N, int K) {
  return (long long)seb200_wgrad_splits(M, N, K) * ((long long)N * K + N);
}

// a: the forward GEMM's descriptor (loader, a[], lda, ln_*, conv geometry, M, K; the weight / output fields are ignored);
// g_out: gradient of the forward GEMM's output rows [M, N] with row stride ldg (for the sub-pixel conv: the [B*T*Fout, 128] view of the
// interleaved output's gradient);  dw: destination of dW through the index map (n, k) -> n * sn + (k / n1) * s0 + (k % n1) * s1 for
// k < k_logical (K may be padded);  db: [N] or NULL.
extern "C" int seb200_wgrad(const SebGemm* a, const float* g_out, long long ldg, int N, int k_logical, int n1, long long sn, long long s0, long long s1,
                            float* dw, float* db, float* workspace, long long workspace_floats, void* stream) {
  SEB_REQUIRE(a && g_out && dw && workspace, SEB_EINVAL, "wgrad: null argument");
  SEB_REQUIRE(a->M > 0 && N > 0 && N % 64 == 0 && N <= 256 && a->K > 0 && a->K % 64 == 0 && k_logical > 0 && k_logical <= a->K && n1 > 0, SEB_EINVAL,
              "wgrad: bad sizes M=%d N=%d K=%d", a->M, N, a->K);
  SEB_REQUIRE(aligned16(g_out) && ldg % 4 == 0 && aligned16(workspace) && a->a[0] && aligned16(a->a[0]), SEB_EALIGN, "wgrad: unaligned operand");
  if (a->loader == SEB_LOAD_ROWS || a->loader == SEB_LOAD_ROWS_LN) SEB_REQUIRE(a->lda % 4 == 0 && a->lda >= a->K, SEB_EALIGN, "wgrad: bad lda");
  if (a->loader == SEB_LOAD_ROWS_LN) SEB_REQUIRE(a->K == 64 && a->ln_gamma && a->ln_beta, SEB_EINVAL, "wgrad: LayerNorm loader needs K == 64 and gamma / beta");
  if (a->loader == SEB_LOAD_CONV) {
    SEB_REQUIRE(a->nslots >= 1 && a->nslots <= 4 && (a->taps_t == 1 || a->taps_t == 2) && a->stride_f >= 1 && a->dil >= 1 &&
                a->K == a->taps_t * 3 * a->nslots * 64 && (long long)a->B * a->T * a->Fout == a->M, SEB_EINVAL, "wgrad: bad conv geometry");
    for (int i = 0; i < a->nslots; ++i) SEB_REQUIRE(a->a[i] && aligned16(a->a[i]), SEB_EALIGN, "wgrad: conv slot %d null/unaligned", i);
  }
  const int S = seb200_wgrad_splits(a->M, N, a->K);
  SEB_REQUIRE(workspace_floats >= (long long)S * ((long long)N * a->K + N), SEB_EINVAL, "wgrad: workspace too small");
  const int rows_per_split = (((a->M + S - 1) / S) + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  const GemmArgs g = wg_args(a);
  float* partial = workspace;
  float* partial_b = db ? workspace + (long long)S * N * a->K : nullptr;
  dim3 grid(S, a->K / 64, N / 64);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (a->loader) {
    case SEB_LOAD_ROWS:    wgrad_kernel<SEB_LOAD_ROWS><<<grid, 256, 0, st>>>(g, g_out, ldg, N, rows_per_split, partial, partial_b); break;
    case SEB_LOAD_ROWS_LN: wgrad_kernel<SEB_LOAD_ROWS_LN><<<grid, 256, 0, st>>>(g, g_out, ldg, N, rows_per_split, partial, partial_b); break;
    case SEB_LOAD_CONV:    wgrad_kernel<SEB_LOAD_CONV><<<grid, 256, 0, st>>>(g, g_out, ldg, N, rows_per_split, partial, partial_b); break;
    default: set_error("wgrad: loader %d is not supported", a->loader); return SEB_EUNSUPPORTED;
  }
  SEB_CHECK_LAUNCH("wgrad_kernel");
  const long long total = (long long)N * a->K + (db ? N : 0);
  wgrad_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, partial_b, S, N, a->K, k_logical, n1, sn, s0, s1, dw, db);
  SEB_CHECK_LAUNCH("wgrad_finish_kernel");
  return 0;
}
