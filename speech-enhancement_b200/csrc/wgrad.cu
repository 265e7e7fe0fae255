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
// reduces rows [s * rows_per_split, ...) of the 64 x 64 tile (n-tile nt, K chunk kc) on the tensor cores (3xTF32 mma.sync, fp32-grade;
// operands staged 32 rows at a time in shared memory) and writes partial[s][n][k]; seb200_wgrad_finish sums the S partials in a
// fixed order (deterministic, no atomics) and scatters them through a two-level index map into the parameter's own layout (conv weights
// are [Cout, Cin, kt, kf] while K runs (tap, slot, channel)) -- typically straight into the flat gradient buffer the all-reduce sends.
#include "gemm_engine.cuh"
#include "mma_tf32.cuh"
#include <stdlib.h>

namespace seb {

constexpr int WG_ROWS = 32;          // rows staged per step
constexpr int WG_LD = 72;            // padded row length (words): the fragment reads (4 rows x 8 columns per instruction) hit 32 distinct banks

// The 64 x 64 output tile is owned by 8 warps in a 2 (n) x 4 (k) grid, 32 x 16 outputs each, on mma.sync m16n8k8 TF32 with the 3xTF32 split
// (mma_tf32.cuh): A = G^T (rows n, contraction over the staged rows m), B = the layer input (rows m, columns k).  Both operands are split
// into (hi, lo) TF32 planes ONCE when they are staged (every element is read by two or four warps), so the main loop is pure LDS + MMA.
// Every 32-row step accumulates in a fresh tensor-core accumulator that is then added to the running sum with IEEE fp32 adds: the long
// reduction over 10^4 .. 10^5 rows never sits inside the tensor pipe's accumulator.
// KC = 64-wide K chunks per CTA, NC = 64-wide n-tiles per CTA.  With KC = 2 (every K that is a multiple of 128: the convolutions, the feed-forward's
// second layer) the staged G rows serve twice the tensor work: half the G loads / splits / stores and half the block barriers per MMA; with NC = 2
// (K = 64 and N a multiple of 128: q|k|v, the feed-forward's first layer, the GLU pointwise conv) the same holds for the staged input rows --
// whose fetch includes the fused LayerNorm.
// Two staging sets: the (hi, lo) planes of step i + 1 are written while the MMAs of step i read the other set -- one block barrier per step and the
// split / store work overlaps the tensor work of the CTA's other warps (one set: two barriers per step, tensor pipe 30 - 36 % busy, profiles/r2/ncu_full_wgrad.txt).
template <int KC, int NC> constexpr int wg_stage_rows() { return (2 * NC + 2 * KC) * WG_ROWS; }
template <int KC, int NC> constexpr int wg_smem_bytes() { return 2 * wg_stage_rows<KC, NC>() * WG_LD * 4; }

template <int LK, int KC, int NC>
__global__ void __launch_bounds__(256) wgrad_kernel(const GemmArgs g, const float* __restrict__ G, long long ldg, int N, int rows_per_split,
                                                    float* __restrict__ partial, float* __restrict__ partial_b) {
  static_assert(KC * NC <= 2, "one operand may be doubled");
  extern __shared__ __align__(16) uint32_t wg_sm[];
  typedef uint32_t (*Plane)[WG_LD];
  Plane Gh[NC], Gl[NC], Ah[KC], Al[KC];
  {
    Plane p0 = reinterpret_cast<Plane>(wg_sm);
#pragma unroll
    for (int j = 0; j < NC; ++j) { Gh[j] = p0 + (2 * j) * WG_ROWS; Gl[j] = p0 + (2 * j + 1) * WG_ROWS; }
#pragma unroll
    for (int j = 0; j < KC; ++j) { Ah[j] = p0 + (2 * NC + 2 * j) * WG_ROWS; Al[j] = p0 + (2 * NC + 2 * j + 1) * WG_ROWS; }
  }
  const int tid = threadIdx.x, sub = tid & 7, rloc = tid >> 3;      // staging: 8 lanes per row, 8 floats each
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int wn = (warp >> 2) * 32, wk = (warp & 3) * 16;            // this warp's 32 x 16 block of every 64 x 64 tile
  const int split = blockIdx.x, kc = blockIdx.y * KC, n0 = blockIdx.z * (64 * NC);
  const int m_lo = split * rows_per_split;
  const int m_hi = min(g.M, m_lo + rows_per_split);
  float acc[NC][KC][2][2][4];
#pragma unroll
  for (int nn = 0; nn < NC; ++nn)
#pragma unroll
    for (int q = 0; q < KC; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[nn][q][i][j][e] = 0.f;
  float bsum = 0.f;
  const bool want_bias = partial_b != nullptr && kc == 0;

  // software pipeline: the global loads of step i + 1 are issued before the MMAs of step i, so their latency hides behind the tensor work
  float v[KC][8], gv[NC][8];
  auto fetch = [&](int m0) {
    const int m = m0 + rloc;
    typename Loader<LK>::Row row;
    Loader<LK>::init_row(g, (m0 < m_hi && m < m_hi) ? m : g.M, row);       // rows past the split read as zeros
#pragma unroll
    for (int j = 0; j < KC; ++j) Loader<LK>::load(g, row, kc + j, sub, v[j]);
#pragma unroll
    for (int nn = 0; nn < NC; ++nn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) gv[nn][i] = 0.f;
      if (m0 < m_hi && m < m_hi && n0 + nn * 64 + sub * 8 < N) {
        const float* gp = G + (long long)m * ldg + n0 + nn * 64 + sub * 8;
        const float4 g0 = ldg4(gp), g1 = ldg4(gp + 4);
        gv[nn][0] = g0.x; gv[nn][1] = g0.y; gv[nn][2] = g0.z; gv[nn][3] = g0.w; gv[nn][4] = g1.x; gv[nn][5] = g1.y; gv[nn][6] = g1.z; gv[nn][7] = g1.w;
      }
    }
  };
  constexpr int STG = wg_stage_rows<KC, NC>();
  auto store_stage = [&](int so) {                                   // split the fetched rows into (hi, lo) TF32 planes of the staging set at row offset so
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < KC; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tf32::split(v[j][i], h[i], l[i]);
      *reinterpret_cast<uint4*>(&Ah[j][so + rloc][sub * 8]) = make_uint4(h[0], h[1], h[2], h[3]);       // 16-byte stores: conflict-free per quarter warp
      *reinterpret_cast<uint4*>(&Ah[j][so + rloc][sub * 8 + 4]) = make_uint4(h[4], h[5], h[6], h[7]);
      *reinterpret_cast<uint4*>(&Al[j][so + rloc][sub * 8]) = make_uint4(l[0], l[1], l[2], l[3]);
      *reinterpret_cast<uint4*>(&Al[j][so + rloc][sub * 8 + 4]) = make_uint4(l[4], l[5], l[6], l[7]);
    }
#pragma unroll
    for (int nn = 0; nn < NC; ++nn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tf32::split(gv[nn][i], h[i], l[i]);
      *reinterpret_cast<uint4*>(&Gh[nn][so + rloc][sub * 8]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&Gh[nn][so + rloc][sub * 8 + 4]) = make_uint4(h[4], h[5], h[6], h[7]);
      *reinterpret_cast<uint4*>(&Gl[nn][so + rloc][sub * 8]) = make_uint4(l[0], l[1], l[2], l[3]);
      *reinterpret_cast<uint4*>(&Gl[nn][so + rloc][sub * 8 + 4]) = make_uint4(l[4], l[5], l[6], l[7]);
    }
  };
  fetch(m_lo);
  store_stage(0);
  __syncthreads();
  int so = 0;                                                        // row offset of the staging set the MMAs of this step read
  for (int m0 = m_lo; m0 < m_hi; m0 += WG_ROWS, so ^= STG) {
    fetch(m0 + WG_ROWS);                                             // next step's operands (all lanes call it: the LayerNorm loader shuffles); in flight across the MMAs
    float c[NC][KC][2][2][4];
#pragma unroll
    for (int nn = 0; nn < NC; ++nn)
#pragma unroll
      for (int q = 0; q < KC; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) c[nn][q][i][j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
      const int r0 = so + ks * 8 + tq, r1 = r0 + 4;
      uint32_t ah[NC][2][4], al[NC][2][4];
#pragma unroll
      for (int nn = 0; nn < NC; ++nn)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int c0 = wn + mt * 16 + gq;
          ah[nn][mt][0] = Gh[nn][r0][c0]; ah[nn][mt][1] = Gh[nn][r0][c0 + 8]; ah[nn][mt][2] = Gh[nn][r1][c0]; ah[nn][mt][3] = Gh[nn][r1][c0 + 8];
          al[nn][mt][0] = Gl[nn][r0][c0]; al[nn][mt][1] = Gl[nn][r0][c0 + 8]; al[nn][mt][2] = Gl[nn][r1][c0]; al[nn][mt][3] = Gl[nn][r1][c0 + 8];
        }
#pragma unroll
      for (int q = 0; q < KC; ++q)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int c0 = wk + nt * 8 + gq;
          const uint32_t bh[2] = {Ah[q][r0][c0], Ah[q][r1][c0]}, bl[2] = {Al[q][r0][c0], Al[q][r1][c0]};
#pragma unroll
          for (int nn = 0; nn < NC; ++nn)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              tf32::mma(c[nn][q][mt][nt], al[nn][mt], bh);
              tf32::mma(c[nn][q][mt][nt], ah[nn][mt], bl);
              tf32::mma(c[nn][q][mt][nt], ah[nn][mt], bh);
            }
        }
    }
#pragma unroll
    for (int nn = 0; nn < NC; ++nn)
#pragma unroll
      for (int q = 0; q < KC; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nn][q][i][j][e] += c[nn][q][i][j][e];
    if (want_bias && tid < 64 * NC) {
      const int col = tid & 63;
      const Plane gh = (tid >> 6) ? Gh[NC - 1] : Gh[0], gl = (tid >> 6) ? Gl[NC - 1] : Gl[0];      // selects, not a dynamic index into the pointer arrays
      float b = 0.f;
#pragma unroll 8
      for (int r = 0; r < WG_ROWS; ++r) b += __uint_as_float(gh[so + r][col]) + __uint_as_float(gl[so + r][col]);
      bsum += b;
    }
    store_stage(so ^ STG);                                           // the other set: its readers finished before the previous barrier
    __syncthreads();
  }
  const int K = g.K;
#pragma unroll
  for (int nn = 0; nn < NC; ++nn)
#pragma unroll
    for (int q = 0; q < KC; ++q)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const int n = n0 + nn * 64 + wn + mt * 16 + gq + ((e >> 1) << 3);
            const int k = (kc + q) * 64 + wk + nt * 8 + tq * 2;
            if (n < N) *reinterpret_cast<float2*>(partial + ((long long)split * N + n) * K + k) = make_float2(acc[nn][q][mt][nt][e], acc[nn][q][mt][nt][e + 1]);
          }
  if (want_bias && tid < 64 * NC && n0 + tid < N) partial_b[(long long)split * N + n0 + tid] = bsum;
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

// SEB200_WGRAD_KC1=1 keeps one K chunk per CTA everywhere (A/B measurements of the two-chunk form)
static int wg_kc(int K) {
  static const bool kc1 = getenv("SEB200_WGRAD_KC1") && atoi(getenv("SEB200_WGRAD_KC1")) != 0;
  return (!kc1 && K % 128 == 0) ? 2 : 1;
}
static int wg_nc(int N, int K) {
  static const bool kc1 = getenv("SEB200_WGRAD_KC1") && atoi(getenv("SEB200_WGRAD_KC1")) != 0;
  return (!kc1 && wg_kc(K) == 1 && N % 128 == 0) ? 2 : 1;
}

// Number of row splits seb200_wgrad uses for (M, N, K): enough CTAs for two waves of 148 SMs, at least 256 rows per split.
extern "C" int seb200_wgrad_splits(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int tiles = (K / (64 * wg_kc(K))) * (((N + 63) / 64) / wg_nc(N, K));      // two K chunks or two n-tiles per CTA when the shape allows
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
  const int KC = wg_kc(a->K), NC = wg_nc(N, a->K);
  dim3 grid(S, a->K / (64 * KC), N / (64 * NC));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_ROWS, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<2, 1>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_CONV, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<2, 1>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_ROWS, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<1, 2>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_ROWS_LN, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<1, 2>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_CONV, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<1, 2>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_ROWS, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<1, 1>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_ROWS_LN, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<1, 1>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_kernel<SEB_LOAD_CONV, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem_bytes<1, 1>());
    if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
#define SEB_WG_LAUNCH(LK, KCV, NCV) wgrad_kernel<LK, KCV, NCV><<<grid, 256, wg_smem_bytes<KCV, NCV>(), st>>>(g, g_out, ldg, N, rows_per_split, partial, partial_b)
  switch (a->loader) {
    case SEB_LOAD_ROWS:
      if (KC == 2) SEB_WG_LAUNCH(SEB_LOAD_ROWS, 2, 1); else if (NC == 2) SEB_WG_LAUNCH(SEB_LOAD_ROWS, 1, 2); else SEB_WG_LAUNCH(SEB_LOAD_ROWS, 1, 1);
      break;
    case SEB_LOAD_ROWS_LN:
      if (NC == 2) SEB_WG_LAUNCH(SEB_LOAD_ROWS_LN, 1, 2); else SEB_WG_LAUNCH(SEB_LOAD_ROWS_LN, 1, 1);
      break;
    case SEB_LOAD_CONV:
      if (KC == 2) SEB_WG_LAUNCH(SEB_LOAD_CONV, 2, 1); else if (NC == 2) SEB_WG_LAUNCH(SEB_LOAD_CONV, 1, 2); else SEB_WG_LAUNCH(SEB_LOAD_CONV, 1, 1);
      break;
    default: set_error("wgrad: loader %d is not supported", a->loader); return SEB_EUNSUPPORTED;
  }
#undef SEB_WG_LAUNCH
  SEB_CHECK_LAUNCH("wgrad_kernel");
  const long long total = (long long)N * a->K + (db ? N : 0);
  wgrad_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, partial_b, S, N, a->K, k_logical, n1, sn, s0, s1, dw, db);
  SEB_CHECK_LAUNCH("wgrad_finish_kernel");
  return 0;
}
