// attention.cu -- multi-head self-attention core with Shaw relative positions (conformer.py:103-122), flash style:
// the (S, h, n, n) score / rel-pos / probability tensors of the reference are never materialised.
//
//   dots[i, j] = scale * q_i . (k_j + E[clamp(i - j, -512, 512) + 512]),   out_i = softmax_j(dots) . v
//
// variant 0: tensor-core kernel (mma.sync m16n8k16 FP16 operands, fp32 accumulate, online softmax).  The rel-pos term
//            of a 32 x 64 score tile is a second small GEMM  R = Q . E_window^T  (96 distinct offsets) followed by a
//            skewed read  S[i, j] += R[i, i - j - dlo]  through a warp-private shared-memory staging buffer.
//            A 10-bit mantissa is sufficient for the three attention contractions.
// variant 1: one-thread-per-query fp32 kernel; slow, used by the tests to cross-check variant 0.
#include "gemm_engine.cuh"   // ptx:: mbarrier helpers
#include <cuda_fp16.h>
#include <type_traits>

namespace seb {

constexpr int AT_H = 4, AT_D = 16, AT_ROW = 192;   // qkv row = (q | k | v) x (4 heads x 16)
constexpr int AT_MAXPOS = 512;

__device__ __forceinline__ long long seq_base(const SebSeq& sq, int seq) {
  return (long long)(seq / sq.inner) * sq.outer_stride + (seq % sq.inner);
}

// ------------------------------------------------------------------------------------------------
// variant 1: SIMT reference
// ------------------------------------------------------------------------------------------------
constexpr int ATS_BQ = 128, ATS_BK = 64, ATS_LD = 20;

__global__ void __launch_bounds__(ATS_BQ) attention_simt_kernel(const float* __restrict__ qkv, const float* __restrict__ E,
                                                               const SebSeq sq, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int seq = blockIdx.x >> 2, h = blockIdx.x & 3, i0 = blockIdx.y * ATS_BQ;
  const int n = sq.n;
  const long long base = seq_base(sq, seq);
  int e_lo = i0 - (n - 1); if (e_lo < -AT_MAXPOS) e_lo = -AT_MAXPOS;
  int e_hi = i0 + ATS_BQ - 1; if (e_hi > AT_MAXPOS) e_hi = AT_MAXPOS;
  const int erows = e_hi - e_lo + 1;
  float* Es = sm;                       // [erows][20]
  float* Ks = Es + erows * ATS_LD;      // [64][16]
  float* Vs = Ks + ATS_BK * AT_D;       // [64][16]
  for (int idx = threadIdx.x; idx < erows * 4; idx += ATS_BQ) {
    const int r = idx >> 2, part = idx & 3;
    *reinterpret_cast<float4*>(Es + r * ATS_LD + part * 4) = ldg4(E + (long long)(e_lo + AT_MAXPOS + r) * AT_D + part * 4);
  }
  int i = i0 + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;
  float q[AT_D], o[AT_D];
  {
    const float* qp = qkv + (base + (long long)i * sq.pos_stride) * AT_ROW + h * AT_D;
#pragma unroll
    for (int d = 0; d < AT_D; d += 4) { float4 v = ldg4(qp + d); q[d] = v.x * 0.25f; q[d + 1] = v.y * 0.25f; q[d + 2] = v.z * 0.25f; q[d + 3] = v.w * 0.25f; }
#pragma unroll
    for (int d = 0; d < AT_D; ++d) o[d] = 0.f;
  }
  float m = -1e30f, l = 0.f;
  for (int j0 = 0; j0 < n; j0 += ATS_BK) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < ATS_BK * 4; idx += ATS_BQ) {
      const int key = idx >> 2, part = idx & 3, j = j0 + key;
      float4 kv = make_float4(0, 0, 0, 0), vv = kv;
      if (j < n) {
        const float* p = qkv + (base + (long long)j * sq.pos_stride) * AT_ROW + h * AT_D + part * 4;
        kv = ldg4(p + 64); vv = ldg4(p + 128);
      }
      *reinterpret_cast<float4*>(Ks + key * AT_D + part * 4) = kv;
      *reinterpret_cast<float4*>(Vs + key * AT_D + part * 4) = vv;
    }
    __syncthreads();
    const int jn = (n - j0 < ATS_BK) ? n - j0 : ATS_BK;
    for (int jj = 0; jj < jn; ++jj) {
      int d = i - (j0 + jj);
      d = d < -AT_MAXPOS ? -AT_MAXPOS : (d > AT_MAXPOS ? AT_MAXPOS : d);
      const float* er = Es + (d - e_lo) * ATS_LD;
      const float* kr = Ks + jj * AT_D;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < AT_D; ++c) s = fmaf(q[c], kr[c] + er[c], s);
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      l = l * corr + p;
      const float* vr = Vs + jj * AT_D;
#pragma unroll
      for (int c = 0; c < AT_D; ++c) o[c] = fmaf(p, vr[c], o[c] * corr);
      m = mn;
    }
  }
  if (live) {
    const float inv = 1.0f / l;
    float* op = out + (base + (long long)i * sq.pos_stride) * 64 + h * AT_D;
#pragma unroll
    for (int d = 0; d < AT_D; d += 4) st4(op + d, make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv));
  }
}

// ------------------------------------------------------------------------------------------------
// variant 0: FP16 mma.sync (m16n8k16, fp32 accumulate) flash attention with the rel-pos GEMM + skew
//
// Input is the fp16 projection [tokens, 192] = (q * dim_head^-0.5 * log2 e | k | v) written by the QKV GEMM epilogue.
// CTA = 4 warps; a warp owns 32 query rows (two 16-row MMA tiles that share every B fragment) of one (sequence,
// head).  K/V tiles stream through a cp.async double buffer (no register staging, no conversion); B fragments come
// from ldmatrix (.trans for V); a ones column appended to V makes the P V product deliver the softmax row sums.  Per 64-key tile: S = Q K^T (16 MMAs), R = Q E_window^T over the 96 distinct offsets of a 32 x 64 tile
// (20 MMAs, E fragments straight from the fp16 table through L1), skew-add through a warp-private fp32 staging
// tile, online softmax with ex2.approx, O += P V (16 MMAs, V staged transposed).  fp16 operands carry the same
// 10-bit mantissa as TF32 (sufficient: SURVEY appendix B.1b); q is pre-scaled so logits stay far inside fp16 range.
// The grid is flat with the query block as the fastest index, so the CTAs that share a sequence's K/V run together
// and the re-reads hit L2.
// ------------------------------------------------------------------------------------------------
constexpr int A2_BK = 64, A2_LD = 24 /*halfs per K / V row (16 + pad: conflict-free ldmatrix)*/, A2_RLD = 104 /*floats*/, A2_WROWS = 32;
constexpr int A2_TILE_H = A2_BK * A2_LD;                                   // halfs per K (or V) tile
constexpr int A2_STAGES = 3;
constexpr int A2_SMEM = A2_STAGES * 2 * A2_TILE_H * 2 + 4 * A2_WROWS * A2_RLD * 4;   // K,V ring + R staging
constexpr int AT_ROWH = 192;                                               // fp16 qkv row: q(64) | k(64) | v(64)

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D = A . B + C with C in its own registers (an accumulator that starts from a per-row constant needs no copies)
__device__ __forceinline__ void mma_f16_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%11,%11};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
}
// fp16-accumulating variant: the two D registers are half2 (row g: cols 2t, 2t+1 | row g + 8: same cols)
__device__ __forceinline__ void mma_f16_h(uint32_t& d0, uint32_t& d1, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%8};"
               : "=r"(d0), "=r"(d1)
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(0u));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}

// BAND = true (long sequences, n > ~1100): key tiles whose offsets i - j all lie beyond +-512 see ONE embedding row
// (conformer.py:108 clamps the distance), so their rel-pos logit is a per-query constant q_i . E[0] or q_i . E[1024]:
// the score accumulators start from that constant and the R GEMM + skew are skipped (78 % of the tiles at T = 4801).

// ------------------------------------------------------------------------------------------------
// variant 0 (v4): the tensor-core kernel, organised around what ncu showed binds it -- the LSU / shared
// memory pipe (72 % busy with the fp32 skew round trip, 27 % of it bank conflicts), dead warps (641 rows = 20 full
// 32-row blocks + 1 row -> a quarter of the last CTA idles), a 64-key tile spent on ONE live key, and ~11 instructions
// per score.
//   * a warp's 32 query rows are split by parity: MMA tile 0 holds the even rows, tile 1 the odd rows.  Then the rel-pos
//     term of tile mt is  S[rho, c] += R[rho, dd = 2 rho - c + 63]  with R = Q_mt . E[base_mt + dd]^T, base_1 = base_0 + 1:
//     for even c the pair (c, c + 1) sits at an even position of a DESCENDING fp16 row, so the skewed operand is read back
//     as aligned half2 words that ARE the A fragment of an MMA.
//   * the skew-add itself runs on the tensor pipe:  S += A_skew . I  (two identity-selecting B fragments), so the 64
//     scalar LDS + 64 FADD per thread become 32 LDS.32 + 16 HMMA, and the staging tile is fp16 (half the bytes).
//   * lazy running maximum: the accumulators start from -m (the row's reference maximum), exp2 is applied directly, and the
//     reference only moves (with the usual rescale) when a tile exceeds it by more than 2^8 -- the per-score subtraction
//     and the per-tile rescale of O disappear from the common path.
//   * the fp16 embedding table arrives in FRAGMENT ORDER (halfs k = 0,1,8,9, 2,3,10,11, ... of a row; ops.pack_rel_pos),
//     so one 8-byte load per lane fetches a whole B fragment of an offset tile (half the LSU wavefronts of two 4-byte loads);
//   * in-band tiles (all offsets inside +-512) address E with immediates; CTAs have 3 or 4 warps, whichever wastes fewer;
//     a final tile with <= 16 (33..48) live keys runs a 16-key (48-key) body.
// ------------------------------------------------------------------------------------------------
constexpr int A4_PW = 40;                                   // staged R is TRANSPOSED: word (w, row) at w * 40 + row, w < 48, row < 32 -- with the
                                                            // fragment lane maps both the half2 stores (-40 t + g) and the A-fragment reads (25 g + 8 t mod 32) hit 32 distinct banks
constexpr int A4_RW = 48 * A4_PW;                           // words per warp
constexpr float A4_LAZY = 8.0f;
constexpr int A4_STAGES = 4;                                // K/V ring: prefetch distance 2, so a warp may lag the fastest one by a full tile
__host__ __device__ constexpr int a4_smem(int warps) { return A4_STAGES * 2 * A2_TILE_H * 2 + warps * A4_RW * 4; }

// MINB = CTAs per SM the register budget is sized for: 3 (168 registers) for the time axis; 4 (128 registers, a few spills) for
// the 4-warp CTAs of short sequences (frequency axis), where 16 instead of 12 resident warps measured 6 % faster.
template <bool BAND, int MINB>
__global__ void __launch_bounds__(128, MINB) attention_v4_kernel(const __half* __restrict__ qkvh, const __half* __restrict__ Eh,
                                                             const SebSeq sq, int nqb, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ uint64_t full_bar[A4_STAGES], empty_bar[A4_STAGES];
  __half* KV = reinterpret_cast<__half*>(smraw);                       // [stage][K | V][64][24]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  uint32_t* Rw = reinterpret_cast<uint32_t*>(KV + A4_STAGES * 2 * A2_TILE_H) + warp * A4_RW;   // [2 tiles][16 rows][A4_PW] half2 words
  const int sh = blockIdx.x / nqb, qb = blockIdx.x - sh * nqb;
  const int seq = sh >> 2, h = sh & 3;
  const int n = sq.n;
  const long long base = seq_base(sq, seq);
  const __half* hbase = qkvh + h * AT_D;
  const int iw = (qb * nwarps + warp) * A2_WROWS;          // first query row of this warp
  const bool warp_live = iw < n;

  if (tid == 0) {
    for (int i = 0; i < A4_STAGES; ++i) { ptx::mbar_init(&full_bar[i], nthr); ptx::mbar_init(&empty_bar[i], nwarps); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  // K/V tile loader: 64 keys x (K lo, K hi, V lo, V hi) 16-byte chunks = 256 cp.async per tile, <= 3 per thread.  The
  // source pointers are set up once and advance by a constant per tile (keys past the sequence: zero fill from row 0).
  const long long tile_stride_h = (long long)A2_BK * sq.pos_stride * AT_ROWH;        // halfs per key tile
  const __half* seq0 = hbase + base * AT_ROWH;                                         // row 0 of this (sequence, head)
  const __half* ld_src[3];
  uint32_t ld_dst[3];
  int ld_key[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int idx = tid + r * nthr;
    const int key = idx >> 2, c = idx & 3;
    ld_key[r] = idx < 256 ? key : (1 << 30);
    ld_src[r] = seq0 + (long long)key * sq.pos_stride * AT_ROWH + 64 + (c >> 1) * 64 + (c & 1) * 8;
    ld_dst[r] = ptx::smem_u32(KV + (c >> 1) * A2_TILE_H + key * A2_LD + (c & 1) * 8);
  }
  auto issue_tile = [&](int tile) {
    const int stage = tile % A4_STAGES;
    const uint32_t soff = (uint32_t)(stage * 2 * A2_TILE_H * 2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (ld_key[r] < A2_BK) {
        const bool ok = tile * A2_BK + ld_key[r] < n;
        const __half* src = ok ? ld_src[r] + tile * tile_stride_h : seq0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ld_dst[r] + soff), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(&full_bar[stage])) : "memory");
  };
  const int ntiles = (n + A2_BK - 1) / A2_BK;
  issue_tile(0);
  if (ntiles > 1) issue_tile(1);

  // Q fragments: tile mt holds rows iw + 2 rho + mt, rho = g and g + 8
  uint32_t qa[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    int r0 = iw + 2 * g + mt, r1 = r0 + 16;
    r0 = r0 < n ? r0 : n - 1;
    r1 = r1 < n ? r1 : n - 1;
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(hbase + (base + (long long)r0 * sq.pos_stride) * AT_ROWH) + t;
    const uint32_t* q1 = reinterpret_cast<const uint32_t*>(hbase + (base + (long long)r1 * sq.pos_stride) * AT_ROWH) + t;
    qa[mt][0] = __ldg(q0); qa[mt][1] = __ldg(q1); qa[mt][2] = __ldg(q0 + 4); qa[mt][3] = __ldg(q1 + 4);
  }
  float o[2][3][4];
  float mrow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    mrow[mt][0] = mrow[mt][1] = 0.f;
#pragma unroll
    for (int x = 0; x < 3; ++x)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][x][e] = 0.f;
  }
  const uint32_t ones = (g == 0) ? 0x3C003C00u : 0u;
  // identity-selecting B fragment: B[kk][nn] = (kk == nn); this thread holds nn = g, kk = 2t, 2t + 1
  const uint32_t idf = (2 * t == g) ? 0x00003C00u : ((2 * t + 1 == g) ? 0x3C000000u : 0u);
  float c_far[2][2][2];        // [far side][mt][row g / g+8]
  if (BAND) {
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const uint2 ef = __ldg(reinterpret_cast<const uint2*>(Eh + side * (2 * AT_MAXPOS) * AT_D) + t);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
        mma_f16(r4, qa[mt], ef.x, ef.y);
        c_far[side][mt][0] = r4[0]; c_far[side][mt][1] = r4[2];
      }
    }
  }
  const int lm_i = lane >> 3, lm_r = lane & 7;
  const int k_off = ((lm_i >> 1) * 8 + lm_r) * A2_LD + (lm_i & 1) * 8;
  const int v_off = ((lm_i & 1) * 8 + lm_r) * A2_LD + (lm_i >> 1) * 8;
  // staged-R addressing (words): write (row, dd pair) -> (44 - 4 nt + t) * PW + row ; read A fragment -> (16 - rho + 8 kb + t (+ 4)) * PW + row
  uint32_t* rw_lo = Rw + (44 + t) * A4_PW + g;             // + mt * 16, - 4 nt * PW ; rows g + 8: + 8
  const uint32_t* rr = Rw + (16 - g + t) * A4_PW + g;      // + mt * 16 + 8 kb * PW ; rows g + 8: + 8 - 8 * PW

  // one key tile of NT n-tiles (8: 64 keys; 2: the 16-key tail body)
  auto tile_body = [&](auto nt_tag, const __half* Ks, const __half* Vs, int j0, bool first) {
    constexpr int NT = decltype(nt_tag)::value;
    constexpr int NKB = NT / 2;                             // 16-key blocks
    const int b0 = iw - j0 - 63;                            // offset of column dd = 0 for tile 0 (tile 1: + 1)
    const int far = !BAND ? -1 : (b0 >= AT_MAXPOS ? 1 : (b0 + 96 <= -AT_MAXPOS ? 0 : -1));
    float s[2][NT][4];
    constexpr int NT_LO = 8 - NT;                            // dd = 2 rho - c + 63 with c < 8 NT  ->  dd >= 64 - 8 NT
    const bool rel = !BAND || far < 0;                       // this tile needs the R GEMM + skew (else: far-field constant)
    const bool inband = (b0 >= -AT_MAXPOS) && (b0 + 96 <= AT_MAXPOS);
    // E fragments of tile mt: issued early so that the loads fly under the MMAs in front of their use
    uint2 ef[12];
    auto load_e = [&](int mt) {
      if (inband) {                   // all offsets inside +-512: one base pointer, immediates
        const uint2* ebase = reinterpret_cast<const uint2*>(Eh + (long long)(b0 + mt + 7 - g + AT_MAXPOS) * AT_D) + t;
#pragma unroll
        for (int nt = NT_LO; nt < 12; ++nt) ef[nt] = __ldg(ebase + nt * 8 * (AT_D / 4));
      } else {                        // tile straddles the clamp (conformer.py:108)
#pragma unroll
        for (int nt = NT_LO; nt < 12; ++nt) {
          int d = b0 + mt + nt * 8 + 7 - g;
          d = d < -AT_MAXPOS ? -AT_MAXPOS : (d > AT_MAXPOS ? AT_MAXPOS : d);
          ef[nt] = __ldg(reinterpret_cast<const uint2*>(Eh + (d + AT_MAXPOS) * AT_D) + t);
        }
      }
    };
    // R[rho, dd] = q . E[clamp(b0 + mt + dd)], dd in [0, 96) -> fp16, transposed, descending (rows < 8 use n-tiles <= 9, rows >= 8
    // n-tiles >= 2).  Column n of offset tile nt is dd = 8 nt + 7 - n (lane g fetched that row), so the half2 the fp16-accumulating
    // MMA returns for columns (2t, 2t + 1) already is one word of the descending staging row: no conversion, no packing.
    auto r_gemm = [&](int mt) {
#pragma unroll
      for (int nt = NT_LO; nt < 12; ++nt) {
        uint32_t d0, d1;
        mma_f16_h(d0, d1, qa[mt], ef[nt].x, ef[nt].y);
        if (nt < 10) rw_lo[mt * 16 - 4 * nt * A4_PW] = d0;
        if (nt >= 2) rw_lo[mt * 16 + 8 - 4 * nt * A4_PW] = d1;
      }
    };
    // skew-add on the tensor pipe: S[:, 16 kb : 16 kb + 16] += A_skew . [I | 0], [0 | I]
    auto skew_add = [&](int mt) {
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const uint32_t* p = rr + mt * 16 + 8 * kb * A4_PW;
        const uint32_t a[4] = {p[0], p[8 - 8 * A4_PW], p[4 * A4_PW], p[8 - 4 * A4_PW]};
        mma_f16(s[mt][2 * kb], a, idf, 0u);
        mma_f16(s[mt][2 * kb + 1], a, 0u, idf);
      }
    };
    if (rel) load_e(0);
    // ---- content scores start from -m (lazy reference maximum) plus the far-field rel-pos constant
    float cinit[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        float c0 = -mrow[mt][rh];
        if (BAND) c0 += (far == 1 ? c_far[1][mt][rh] : (far == 0 ? c_far[0][mt][rh] : 0.f));
        cinit[mt][rh] = c0;
      }
#pragma unroll
    for (int np = 0; np < NKB; ++np) {
      uint32_t kb[4];
      ldsm_x4(kb, Ks + np * 16 * A2_LD + k_off);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_f16_c(s[mt][2 * np], qa[mt], kb[0], kb[1], cinit[mt][0], cinit[mt][1]);
        mma_f16_c(s[mt][2 * np + 1], qa[mt], kb[2], kb[3], cinit[mt][0], cinit[mt][1]);
      }
    }
    if (rel) {
      r_gemm(0);
      asm volatile("" ::: "memory");
      load_e(1);                      // in flight under tile 0's skew-add
      __syncwarp();
      skew_add(0);
      r_gemm(1);
      __syncwarp();
      skew_add(1);
      __syncwarp();
    }
    if (j0 + NT * 8 > n) {   // mask keys beyond the sequence (last tile only)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (j0 + nt * 8 + 2 * t + (e & 1) >= n) s[mt][nt][e] = -1e30f;
    }
    // ---- lazy online softmax (base 2): s already holds logit - m
    float tmax[2][2];
    bool move = first;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        float mx = -1e30f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mx = fmaxf(mx, fmaxf(s[mt][nt][2 * rh], s[mt][nt][2 * rh + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        tmax[mt][rh] = mx;
        move = move || (mx > A4_LAZY);
      }
    if (__any_sync(0xffffffffu, move)) {     // rare after the first tile: move the reference maximum and rescale
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const float delta = first ? tmax[mt][rh] : fmaxf(tmax[mt][rh], 0.f);
          const float corr = first ? 1.0f : ex2_approx(-delta);
          mrow[mt][rh] += delta;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) { s[mt][nt][2 * rh] -= delta; s[mt][nt][2 * rh + 1] -= delta; }
#pragma unroll
          for (int x = 0; x < 3; ++x) { o[mt][x][2 * rh] *= corr; o[mt][x][2 * rh + 1] *= corr; }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[mt][nt][e] = ex2_approx(s[mt][nt][e]);
    // ---- O += P [V | 1]
#pragma unroll
    for (int ks = 0; ks < NKB; ++ks) {
      uint32_t vb[4];
      ldsm_x4_trans(vb, Vs + ks * 16 * A2_LD + v_off);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t pa[4] = {pack_h2(s[mt][2 * ks][0], s[mt][2 * ks][1]), pack_h2(s[mt][2 * ks][2], s[mt][2 * ks][3]),
                                pack_h2(s[mt][2 * ks + 1][0], s[mt][2 * ks + 1][1]), pack_h2(s[mt][2 * ks + 1][2], s[mt][2 * ks + 1][3])};
        mma_f16(o[mt][0], pa, vb[0], vb[1]);
        mma_f16(o[mt][1], pa, vb[2], vb[3]);
        mma_f16(o[mt][2], pa, ones, ones);
      }
    }
  };

  for (int tile = 0; tile < ntiles; ++tile) {
    const int stage = tile % A4_STAGES;
    if (tile + 2 < ntiles) {
      if (tile >= 2) ptx::mbar_wait(&empty_bar[(tile + 2) % A4_STAGES], (uint32_t)((tile - 2) / A4_STAGES) & 1u);
      issue_tile(tile + 2);
    }
    ptx::mbar_wait(&full_bar[stage], (uint32_t)(tile / A4_STAGES) & 1u);
    const int j0 = tile * A2_BK;
    const __half* Ks = KV + stage * 2 * A2_TILE_H;
    const __half* Vs = Ks + A2_TILE_H;
    if (warp_live) {
      const int rem = n - j0;        // live keys of this tile: a short last tile runs a narrower body (16 / 48 keys)
      if (rem <= 16) tile_body(std::integral_constant<int, 2>{}, Ks, Vs, j0, tile == 0);
      else if (rem > 32 && rem <= 48) tile_body(std::integral_constant<int, 6>{}, Ks, Vs, j0, tile == 0);
      else tile_body(std::integral_constant<int, 8>{}, Ks, Vs, j0, tile == 0);
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
  }
  if (!warp_live) return;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const float l = __shfl_sync(0xffffffffu, o[mt][2][2 * rh], lane & ~3);
      const int i = iw + 2 * (g + 8 * rh) + mt;
      if (i < n) {
        const float inv = 1.0f / l;
        float* op = out + (base + (long long)i * sq.pos_stride) * 64 + h * AT_D + 2 * t;
        *reinterpret_cast<float2*>(op) = make_float2(o[mt][0][2 * rh] * inv, o[mt][0][2 * rh + 1] * inv);
        *reinterpret_cast<float2*>(op + 8) = make_float2(o[mt][1][2 * rh] * inv, o[mt][1][2 * rh + 1] * inv);
      }
    }
}

int attention_tc_launch(const __half* qkvh, const __half* Eh, const SebSeq* seq, float* out, cudaStream_t st);   // attention_tc.cu

}  // namespace seb

using namespace seb;

extern "C" int seb200_attention(const void* qkv, const float* rel_pos_emb, const void* rel_pos_emb_h, const SebSeq* seq, float* out, int variant, void* stream) {
  SEB_REQUIRE(qkv && seq && out && aligned16(qkv) && aligned16(out), SEB_EINVAL, "attention: null/unaligned argument");
  if (variant == 1) SEB_REQUIRE(rel_pos_emb && aligned16(rel_pos_emb), SEB_EINVAL, "attention: the fp32 variant needs rel_pos_emb");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0 && seq->nseq <= (1 << 28), SEB_EINVAL, "attention: bad sequence descriptor");
  const int n = seq->n;
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 1) {
    int erows = n + ATS_BQ - 1; if (erows > 2 * AT_MAXPOS + 1) erows = 2 * AT_MAXPOS + 1;
    const int smem = (erows * ATS_LD + 2 * ATS_BK * AT_D) * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(attention_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(seq->nseq * AT_H, (n + ATS_BQ - 1) / ATS_BQ);
    SEB_REQUIRE(grid.y <= 65535u, SEB_EINVAL, "attention: sequence too long");
    attention_simt_kernel<<<grid, ATS_BQ, smem, st>>>(reinterpret_cast<const float*>(qkv), rel_pos_emb, *seq, out);
    SEB_CHECK_LAUNCH("attention_simt_kernel");
    return 0;
  }
  SEB_REQUIRE(rel_pos_emb_h && aligned16(rel_pos_emb_h), SEB_EINVAL, "attention: the tensor-core variant needs the fp16 copy of rel_pos_emb");
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(attention_v4_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, a4_smem(4));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_v4_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, a4_smem(4));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_v4_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, a4_smem(4));
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  if (variant == 3)      // tcgen05 kernel (attention_tc.cu); same inputs as variant 0
    return attention_tc_launch(reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(rel_pos_emb_h), seq, out, st);
  const bool band = n > 2 * AT_MAXPOS + 128;     // far-field shortcut pays once a sizeable share of the (query, key) tiles lies beyond the clamp
  if (variant == 0) {
    // 3 or 4 warps (32 query rows each) per CTA, whichever leaves fewer idle warps in the last CTA of a (sequence, head)
    const int nb = (n + A2_WROWS - 1) / A2_WROWS;
    const int w4 = ((nb + 3) / 4) * 4 - nb, w3 = ((nb + 2) / 3) * 3 - nb;
    const int W = (w3 < w4) ? 3 : 4;
    const int nqb = (nb + W - 1) / W;
    const long long nblocks = (long long)seq->nseq * AT_H * nqb;
    SEB_REQUIRE(nblocks < 2147483647LL, SEB_EINVAL, "attention: grid too large");
    if (band)
      attention_v4_kernel<true, 3><<<(unsigned)nblocks, W * 32, a4_smem(W), st>>>(reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(rel_pos_emb_h), *seq, nqb, out);
    else if (W == 4 && n <= 256)
      attention_v4_kernel<false, 4><<<(unsigned)nblocks, W * 32, a4_smem(W), st>>>(reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(rel_pos_emb_h), *seq, nqb, out);
    else
      attention_v4_kernel<false, 3><<<(unsigned)nblocks, W * 32, a4_smem(W), st>>>(reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(rel_pos_emb_h), *seq, nqb, out);
    SEB_CHECK_LAUNCH("attention_v4_kernel");
    return 0;
  }
  set_error("attention: variant %d is not implemented (0: mma.sync, 1: fp32 SIMT cross-check, 3: tcgen05)", variant);
  return SEB_EUNSUPPORTED;
}
