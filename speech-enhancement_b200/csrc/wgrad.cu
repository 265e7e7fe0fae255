// wgrad.cu -- weight gradients of the generator's dense contractions (training step; loss.backward() at
// /root/reference/core/function.py:274 produces them through ATen's conv / addmm backward kernels).
//
//   dW[n, k] = sum_m G[m, n] * A[m, k]          db[n] = sum_m G[m, n]
//
// G [M, N] is the gradient of the layer's output rows (fp32), A [M, K] is the layer's INPUT exactly as the forward GEMM engine read it
// -- the same Loader<> structs of gemm_engine.cuh (plain rows, rows with LayerNorm fused, implicit-GEMM conv gather over the dense
// block's slot list), so a forward layer and its weight gradient share one description of the operand (SebGemm).
//
// The contraction runs over the M pixels / tokens (10^5 .. 10^7), the output is tiny (N x K <= 256 x 1536): split-M.  CTA (s, tile) reduces rows
// [s * rows_per_split, ...) of one 128 x 64 output tile on tcgen05 (three bf16 planes per operand, fp32-grade; wgrad_tc_kernel below) and writes
// partial[s][n][k]; seb200_wgrad_finish sums the S partials in a fixed order (deterministic, no atomics) and scatters them through a two-level index
// map into the parameter's own layout (conv weights are [Cout, Cin, kt, kf] while K runs (tap, slot, channel)) -- typically straight into the flat
// gradient buffer the all-reduce sends.  (Until round 2, session 3 this contraction ran on 3xTF32 mma.sync: tools/experiments/wgrad_mmasync_3xtf32_kernel.cu.txt,
// 15.9 ms per training step against 10.0 ms.  What bounds the tcgen05 form is shared-memory traffic -- 37 KB of plane stores + 72 KB of operand reads by the
// six products per 32-row step, L1 / shared pipe 57 - 73 % busy -- next to the latency of the loads; a second register set for the loads changed nothing.)
#include "gemm_engine.cuh"
#include <stdlib.h>

namespace seb {

// dW[n, k] = sum_s partial[s][n][k] -> dw[n * sn + (k / n1) * s0 + (k % n1) * s1] for k < k_logical;  db[n] = sum_s partial_b[s][n]
// CTA = 64 elements x 4 split lanes: lane q sums the splits p = q (mod 4), the four lane sums are added in lane order through shared memory (one thread per
// element walking up to 148 splits was a serial chain of dependent-latency loads: 21 us per call, 79 calls per step).  Fixed order: deterministic.
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b, int S, int N, int K,
                                                           int k_logical, int n1, long long sn, long long s0, long long s1,
                                                           float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[4][64];
  const long long total = (long long)N * K;
  const int e = threadIdx.x & 63, q = threadIdx.x >> 6;
  const long long idx = (long long)blockIdx.x * 64 + e;
  const float* src = nullptr;
  long long stride = 0;
  if (idx < total) { src = partial + idx; stride = total; }
  else if (db != nullptr && idx < total + N) { src = partial_b + (idx - total); stride = N; }
  float a0 = 0.f, a1 = 0.f;
  if (src) {
    int p = q;
    for (; p + 4 < S; p += 8) { a0 += src[(long long)p * stride]; a1 += src[(long long)(p + 4) * stride]; }
    if (p < S) a0 += src[(long long)p * stride];
  }
  sm[q][e] = a0 + a1;
  __syncthreads();
  if (q == 0 && src) {
    const float s = (sm[0][e] + sm[1][e]) + (sm[2][e] + sm[3][e]);
    if (idx < total) {
      const int n = (int)(idx / K), k = (int)(idx - (long long)n * K);
      if (k < k_logical) dw[(long long)n * sn + (long long)(k / n1) * s0 + (long long)(k % n1) * s1] = s;
    } else {
      db[(int)(idx - total)] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// The split-M contraction with the tensor work on tcgen05 / tensor memory.  Both operands have the reduction index m as
// their slow index (G [m][n], A [m][k]), i.e. both are MN-MAJOR UMMA operands: a 32-row step is staged as three bf16 planes (hi | mid | lo, 2^-25
// operand error) in the no-swizzle canonical layout (8 rows x 16 bytes core matrices; 8 consecutive features of one row = one 16-byte store), and one
// thread of a ninth warp issues 6 tcgen05.mma (M = 128, K = 16, N = 192 / 128 / 64: the six products of the three-plane split, hi x hi in its own accumulator) per step
// as soon as the eight staging warps have filled a set (full / mma_done mbarriers per set, no block barrier in the loop).  X is the 128-wide side, Y the 64-wide side:
//   MODE 0 (K a multiple of 128, or N = 64):  X = two 64-wide K chunks of the layer input, Y = 64 gradient columns   D[k, n]
//   MODE 1 (K = 64, N a multiple of 128):     X = 128 gradient columns, Y = the layer input                          D[n, k]
// The tensor pipe's fp32 accumulation truncates (section 4b of DESIGN.md): every WT_FLUSH steps (256 rows, 16 MMAs into each accumulator) the
// accumulators are read back and added to a shared-memory fp32 tile with IEEE adds; the split partials go through the same finish kernel as before.
// ------------------------------------------------------------------------------------------------------------------------------------
constexpr int WT_ROWS = 32;
constexpr int WT_SBO = WT_ROWS * 16 + 16;            // bytes between 8-feature blocks (+16: the eight lanes of a row store to different bank groups)
constexpr int WT_XP = 16 * WT_SBO, WT_YP = 8 * WT_SBO;
constexpr int WT_STAGE = 3 * WT_XP + 3 * WT_YP;      // 38016 bytes
constexpr int WT_ACC_LD = 65;                        // fp32 accumulator tile [128][65] (thread = row: conflict-free)
constexpr int WT_FLUSH = 8;
constexpr int WT_SMEM = 2 * WT_STAGE + 128 * WT_ACC_LD * 4 + 128;

__device__ __forceinline__ uint64_t wt_desc(uint32_t smem_addr) {      // no-swizzle MN-major operand: lbo = next 8 rows along K (128 bytes), sbo = next 8 features
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(WT_SBO >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void wt_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int LK, int MODE>
__global__ void __launch_bounds__(288, 2) wgrad_tc_kernel(const GemmArgs g, const float* __restrict__ G, long long ldg, int N, int rows_per_split,
                                                          float* __restrict__ partial, float* __restrict__ partial_b) {
  extern __shared__ __align__(16) uint8_t wt_raw[];
  __shared__ uint64_t full_bar[2], mma_done[2], acc_bar, acc_free;
  __shared__ uint32_t tmem_base_s;
  const uint32_t sm0 = (ptx::smem_u32(wt_raw) + 127u) & ~127u;
  float* accs = reinterpret_cast<float*>(wt_raw + (sm0 - ptx::smem_u32(wt_raw)) + 2 * WT_STAGE);      // [128][WT_ACC_LD]; the bias scratch [32][128] reuses it at the end
  const int tid = threadIdx.x, warp = tid >> 5, sub = tid & 7, rloc = tid >> 3;
  const int split = blockIdx.x;
  const int kc = MODE == 0 ? 2 * blockIdx.y : blockIdx.y;                  // first 64-wide K chunk of this CTA
  const int n0 = blockIdx.z * (MODE == 0 ? 64 : 128);
  const int m_lo = split * rows_per_split;
  const int m_hi = min(g.M, m_lo + rows_per_split);
  const int nsteps = (m_hi - m_lo + WT_ROWS - 1) / WT_ROWS;
  const bool want_bias = partial_b != nullptr && kc == 0;
  if (tid == 0) {
    ptx::mbar_init(&mma_done[0], 1); ptx::mbar_init(&mma_done[1], 1); ptx::mbar_init(&acc_bar, 1);
    ptx::mbar_init(&full_bar[0], 256); ptx::mbar_init(&full_bar[1], 256); ptx::mbar_init(&acc_free, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(&tmem_base_s, 256);
  if (tid < 128) {
#pragma unroll 8
    for (int y = 0; y < 64; ++y) accs[tid * WT_ACC_LD + y] = 0.f;
  }
  // the second X chunk of MODE 0 is empty when K is not a multiple of 128 (K = 64 / 192): its planes stay zero
  const bool x1_live = MODE == 1 || (kc + 1) * 64 < g.K;
  if (!x1_live) {
    for (int i = tid; i < 2 * 3 * 8 * (WT_SBO / 16); i += 288) {
      const int b = i / (3 * 8 * (WT_SBO / 16)), r = i % (3 * 8 * (WT_SBO / 16));
      const int pl = r / (8 * (WT_SBO / 16)), q = r % (8 * (WT_SBO / 16));
      *reinterpret_cast<uint4*>(wt_raw + (sm0 - ptx::smem_u32(wt_raw)) + b * WT_STAGE + pl * WT_XP + 8 * WT_SBO + q * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async_smem();
  }
  float xv[2][8], yv[8];
  float bsum[MODE == 0 ? 1 : 2][8];
#pragma unroll
  for (int j = 0; j < (MODE == 0 ? 1 : 2); ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[j][i] = 0.f;
  auto ld_g = [&](int m, bool ok, int col, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (ok && col < N) {
      const float* gp = G + (long long)m * ldg + col;
      const float4 g0 = ldg4(gp), g1 = ldg4(gp + 4);
      v[0] = g0.x; v[1] = g0.y; v[2] = g0.z; v[3] = g0.w; v[4] = g1.x; v[5] = g1.y; v[6] = g1.z; v[7] = g1.w;
    }
  };
  auto fetch = [&](int m0) {
    const int m = m0 + rloc;
    const bool ok = m0 < m_hi && m < m_hi;
    typename Loader<LK>::Row row;
    Loader<LK>::init_row(g, ok ? m : g.M, row);                            // rows past the split read as zeros
    if (MODE == 0) {
      Loader<LK>::load(g, row, kc, sub, xv[0]);
      if (x1_live) Loader<LK>::load(g, row, kc + 1, sub, xv[1]);
      ld_g(m, ok, n0 + sub * 8, yv);
    } else {
      ld_g(m, ok, n0 + sub * 8, xv[0]);
      ld_g(m, ok, n0 + 64 + sub * 8, xv[1]);
      Loader<LK>::load(g, row, kc, sub, yv);
    }
  };
  auto store8 = [&](uint32_t base, uint32_t plane_bytes, const float (&v)[8]) {      // three bf16 planes of 8 features of row rloc
    uint4 hi, mid, lo;
    split3_bf16x2(v[0], v[1], hi.x, mid.x, lo.x);
    split3_bf16x2(v[2], v[3], hi.y, mid.y, lo.y);
    split3_bf16x2(v[4], v[5], hi.z, mid.z, lo.z);
    split3_bf16x2(v[6], v[7], hi.w, mid.w, lo.w);
    const uint32_t a = base + (uint32_t)((rloc >> 3) * 128 + (rloc & 7) * 16);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + plane_bytes), "r"(mid.x), "r"(mid.y), "r"(mid.z), "r"(mid.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + 2 * plane_bytes), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
  };
  auto store_stage = [&](int b) {
    const uint32_t st = sm0 + (uint32_t)(b * WT_STAGE);
    store8(st + (uint32_t)(sub * WT_SBO), WT_XP, xv[0]);
    if (x1_live) store8(st + (uint32_t)((8 + sub) * WT_SBO), WT_XP, xv[1]);
    store8(st + 3 * WT_XP + (uint32_t)(sub * WT_SBO), WT_YP, yv);
    if (want_bias) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) bsum[0][i] += yv[i];
        else { bsum[0][i] += xv[0][i]; bsum[MODE == 0 ? 0 : 1][i] += xv[1][i]; }
      }
    }
  };
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr uint32_t IDESC_B = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);      // f32 accumulate, bf16 x bf16, both operands MN-major, M = 128
  constexpr uint32_t IDESC_N192 = IDESC_B | ((uint32_t)(192 >> 3) << 17), IDESC_N128 = IDESC_B | ((uint32_t)(128 >> 3) << 17), IDESC_N64 = IDESC_B | ((uint32_t)(64 >> 3) << 17);
  auto is_flush = [&](int i) { return (i % WT_FLUSH == WT_FLUSH - 1) || i == nsteps - 1; };
  if (warp == 8) {
    // ================= MMA issuer (one thread): step i as soon as its staging set is full =================
    if (tid == 256) {
      for (int i = 0; i < nsteps; ++i) {
        const int b = i & 1;
        ptx::mbar_wait(&full_bar[b], (uint32_t)(i >> 1) & 1u);
        if (i > 0 && i % WT_FLUSH == 0) ptx::mbar_wait(&acc_free, (uint32_t)(i / WT_FLUSH - 1) & 1u);      // the previous run's accumulators have been read
        ptx::tc_fence_after();
        const uint32_t st = sm0 + (uint32_t)(b * WT_STAGE);
        const uint32_t first = (i % WT_FLUSH == 0) ? 0u : 1u;
#pragma unroll
        for (int k = 0; k < WT_ROWS / 16; ++k) {
          const uint32_t off = (uint32_t)(k * 256);
          // the three Y planes are consecutive 8-block groups of ONE MN-major operand, so an N = 192 / 128 / 64 instruction multiplies an X plane with
          // [Y_hi | Y_mid | Y_lo] / [Y_hi | Y_mid] / [Y_hi] in one read of that X plane (three instead of six operand reads of X per k-step); the column
          // bases make the equal-magnitude terms land in the same accumulator: A0 = hi.hi | A1 = hi.mid + mid.hi | A2 = hi.lo + mid.mid + lo.hi
          const uint64_t xh = wt_desc(st + off), xm = wt_desc(st + WT_XP + off), xl = wt_desc(st + 2 * WT_XP + off);
          const uint64_t yh = wt_desc(st + 3 * WT_XP + off);
          const uint32_t f = (k == 0) ? first : 1u;
          ptx::mma_bf16(tmem_base, xh, yh, IDESC_N192, f);
          ptx::mma_bf16(tmem_base + 64u, xm, yh, IDESC_N128, 1u);
          ptx::mma_bf16(tmem_base + 128u, xl, yh, IDESC_N64, 1u);
        }
        ptx::tc_commit(&mma_done[b]);
        if (is_flush(i)) ptx::tc_commit(&acc_bar);
      }
    }
  } else {
    // ================= staging warps: fetch -> split -> store, one step ahead of the tensor work; warps 0 - 3 also flush the accumulators =================
    fetch(m_lo);
    int nflush = 0;
    for (int i = 0; i < nsteps; ++i) {
      const int b = i & 1;
      if (i >= 2) ptx::mbar_wait(&mma_done[b], (uint32_t)((i - 2) >> 1) & 1u);          // the MMAs that read this staging set (step i - 2) are done
      store_stage(b);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&full_bar[b]);
      if (i + 1 < nsteps) fetch(m_lo + (i + 1) * WT_ROWS);                    // next step's operands: in flight across the waits below (all lanes call it: the LayerNorm loader shuffles)
      if (is_flush(i)) {
        if (tid < 128) {                                                      // thread = accumulator lane = X feature
          ptx::mbar_wait(&acc_bar, (uint32_t)nflush & 1u);
          ptx::tc_fence_after();
          const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16);
          float* arow = accs + tid * WT_ACC_LD;
#pragma unroll
          for (int c = 0; c < 64; c += 16) {
            float v0[16], v1[16], v2[16];
            wt_ld16(ta + (uint32_t)c, v0);
            wt_ld16(ta + 64u + (uint32_t)c, v1);
            wt_ld16(ta + 128u + (uint32_t)c, v2);
#pragma unroll
            for (int e = 0; e < 16; ++e) arow[c + e] += (v2[e] + v1[e]) + v0[e];
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&acc_free);
        }
        ++nflush;
      }
    }
  }
  __syncthreads();
  // ---- partial tile: MODE 0  D[x = k, y = n] -> partial[split][n0 + y][kc * 64 + x];  MODE 1  D[x = n, y = k] -> partial[split][n0 + x][kc * 64 + y]
  const int K = g.K;
  if (tid < 128) {
    const float* arow = accs + tid * WT_ACC_LD;
    if (MODE == 0) {
      const int k = kc * 64 + tid;
      if (k < K) {
#pragma unroll 8
        for (int y = 0; y < 64; ++y) partial[((long long)split * N + n0 + y) * K + k] = arow[y];
      }
    } else {
      const int n = n0 + tid;
      if (n < N) {
        float* dst = partial + ((long long)split * N + n) * K + kc * 64;
#pragma unroll
        for (int y = 0; y < 64; y += 4) *reinterpret_cast<float4*>(dst + y) = make_float4(arow[y], arow[y + 1], arow[y + 2], arow[y + 3]);
      }
    }
  }
  if (want_bias) {                                                          // column sums of the gradient rows this CTA staged: 32 row partials per column, summed in a fixed order
    __syncthreads();
    float* bs = accs;                                                       // [32][128]
#pragma unroll
    for (int j = 0; j < (MODE == 0 ? 1 : 2); ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) bs[rloc * 128 + j * 64 + sub * 8 + e] = bsum[j][e];
    __syncthreads();
    const int ncol = MODE == 0 ? 64 : 128;
    if (tid < ncol && n0 + tid < N) {
      float t = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) t += bs[r * 128 + tid];
      partial_b[(long long)split * N + n0 + tid] = t;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 256);
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

static int wt_mode(int N, int K) { return (K == 64 && N % 128 == 0) ? 1 : 0; }       // wgrad_tc_kernel: which operand is the 128-wide side

extern "C" int seb200_wgrad_splits(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int tiles = wt_mode(N, K) ? (N / 128) * (K / 64) : ((K + 127) / 128) * ((N + 63) / 64);
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
  const int rows_per_split = (((a->M + S - 1) / S) + WT_ROWS - 1) / WT_ROWS * WT_ROWS;
  const GemmArgs g = wg_args(a);
  float* partial = workspace;
  float* partial_b = db ? workspace + (long long)S * N * a->K : nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    static PerDeviceOnce tc_attr_done;
    if (!tc_attr_done.done()) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<SEB_LOAD_ROWS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<SEB_LOAD_ROWS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<SEB_LOAD_ROWS_LN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<SEB_LOAD_ROWS_LN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<SEB_LOAD_CONV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
      if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
      tc_attr_done.set();
    }
    const int mode = wt_mode(N, a->K);
    dim3 tgrid(S, mode ? a->K / 64 : (a->K + 127) / 128, mode ? N / 128 : N / 64);
#define SEB_WT_LAUNCH(LK, MD) wgrad_tc_kernel<LK, MD><<<tgrid, 288, WT_SMEM, st>>>(g, g_out, ldg, N, rows_per_split, partial, partial_b)
    switch (a->loader) {
      case SEB_LOAD_ROWS: if (mode) SEB_WT_LAUNCH(SEB_LOAD_ROWS, 1); else SEB_WT_LAUNCH(SEB_LOAD_ROWS, 0); break;
      case SEB_LOAD_ROWS_LN: if (mode) SEB_WT_LAUNCH(SEB_LOAD_ROWS_LN, 1); else SEB_WT_LAUNCH(SEB_LOAD_ROWS_LN, 0); break;
      case SEB_LOAD_CONV: SEB_WT_LAUNCH(SEB_LOAD_CONV, 0); break;
      default: set_error("wgrad: loader %d is not supported", a->loader); return SEB_EUNSUPPORTED;
    }
#undef SEB_WT_LAUNCH
    SEB_CHECK_LAUNCH("wgrad_tc_kernel");
  }
  const long long total = (long long)N * a->K + (db ? N : 0);
  wgrad_finish_kernel<<<(unsigned)((total + 63) / 64), 256, 0, st>>>(partial, partial_b, S, N, a->K, k_logical, n1, sn, s0, s1, dw, db);
  SEB_CHECK_LAUNCH("wgrad_finish_kernel");
  return 0;
}
