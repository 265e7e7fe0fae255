// attention_tc.cu -- variant 3 of seb200_attention: the Shaw-relative-position attention core (conformer.py:103-122) on
// tcgen05 tensor cores with every accumulator in tensor memory.
//
// One CTA = (sequence, head PAIR, 64-query block); it walks the keys in tiles of 64.  The 128 accumulator lanes are
// (head-local hl, query parity par, lane l) -> quadrant w = 2 hl + par holds queries i0 + 2 l + par of head 2 hp + hl, so
// every warp is parity-uniform (what lets the skewed rel-pos read address whole fp16 pairs, see below).
//
//   MMA 1a  S[128 x 64]  (fp32) = Aexp[128 x 32] . K[64 x 32]^T     Aexp row = q in its head's 16-wide k slot, zeros in the
//                                                                  other head's slot; K rows are the natural [token, 2 x 16]
//                                                                  slice of the q|k|v projection -> per-head q.k
//   MMA 1b  R[128 x 128] (fp16) = Q[128 x 16] . Ewin[128 x 16]^T    Ewin row c = E[clamp(dtop - c)], dtop = i0 + 63 - j0: all
//                                                                  127 distinct offsets of a 64 x 64 tile, DESCENDING, so
//                                                                  the entries a query needs are an ascending run
//   threads (lane = accumulator row): tcgen05.ld R with pack::16b (half2 words straight from the fp16 accumulator, no
//           conversion) -> 16 x STS.128 into a thread-private shared-memory row -> 33 x LDS.32 at the row's own offset
//           (the per-row skew that registers cannot index) -> FHADD (fp32 += one half of a word, a single instruction on
//           sm_100) onto the content scores from tcgen05.ld S -> lazy running maximum -> ex2 -> fp16 P -> tcgen05.st
//   MMA 3   O[128 x 32]  (fp32) += P[128 x 64] (A operand in TMEM) . V[64 x 32]   V tile in its natural [key, d] order = an
//                                                                  MN-major B operand; row sums are kept by the threads
//
// TMEM (256 columns, two CTAs per SM): S (64) | R (128) | P (32) | O (32).  A fifth warp streams K, V and the E window of
// tile t + 3 with zero-filling cp.async into a 4-slot ring (no-swizzle canonical UMMA layouts, 16-byte chunks placed
// directly) and its lane 0 issues the MMAs.  Four mbarriers: S/R written (tcgen05.commit), S/R read into registers (128
// threads), P stored (128 threads), O updated (tcgen05.commit).  MMA 1 of tile t + 1 is issued as soon as the threads hold
// S(t), R(t) in registers, so the next scores are ready when the threads finish the exponentials of tile t: the softmax
// warps never wait for the tensor pipe in steady state.
#include "gemm_engine.cuh"   // ptx:: mbarrier / tcgen05 helpers
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>

namespace seb {

constexpr int T5_KT = 64, T5_BQ = 64, T5_STAGES = 4, T5_THREADS = 192;
constexpr int T5_ROWH = 192, T5_MAXPOS = 512, T5_D = 16;
constexpr int T5_AEXP = 0, T5_QPL = 8192, T5_STAGE0 = 12288;
constexpr int T5_KS = 0, T5_ES = 4096, T5_VS = 8192, T5_STAGE = 12288;
constexpr int T5_RSCR = T5_STAGE0 + T5_STAGES * T5_STAGE;
constexpr int T5_RPITCH = 272;                      // bytes per private R row: 64 words + 4 -> STS.128 and the skewed LDS.32 are both conflict-free
constexpr int T5_SMEM = T5_RSCR + 128 * T5_RPITCH + 128;
constexpr uint32_t T5_TS = 0, T5_TR = 64, T5_TP = 192, T5_TO = 224, T5_TCOLS = 256;
constexpr float T5_LAZY = 8.0f;

namespace ptx {
// no-swizzle canonical operand: 8-row x 16-byte core matrices; lbo / sbo in bytes (K-major: lbo = next core matrix along K,
// sbo = next 8-row group; MN-major: sbo = next 8 elements along MN, lbo = next 8 rows along K)
__device__ __forceinline__ uint64_t umma_desc_ns(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
#define SEB_R32(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), \
                      "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), \
                      "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), \
                      "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31])
#define SEB_L32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}"
// 32 fp32 columns of this thread's lane
template <int O>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " SEB_L32 ", [%32];" : SEB_R32(r, O) : "r"(taddr) : "memory");
}
// 64 columns holding 16-bit values -> 32 registers (column 2k in the low half of register k)
template <int O>
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " SEB_L32 ", [%32];" : SEB_R32(r, O) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_pack16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
                 "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait5() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Lean bounded wait: try_wait with a suspend-time hint (the warp sleeps in hardware until the phase completes or the hint
// expires) and a 3-instruction retry loop.  The generic ptx::mbar_wait re-reads the clock every iteration; with two
// resident CTAs its ~12-instruction spin took issue slots from the other CTA's working warps (43 % of all issued
// instructions in the first profile).
#ifndef T5_WAIT_HINT_NS
#define T5_WAIT_HINT_NS 4000
#endif
template <int MODE>
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  if (MODE == 0) { mbar_wait(bar, parity); return; }
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    if (MODE == 2 && it) __nanosleep(40);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();     // a protocol bug must surface as a launch error, never as a hung GPU
}
// L1-allocating variant for the embedding window: beyond the +-512 clamp every row of the window is the SAME table row, and
// with .cg all CTAs of a long sequence hammer two L2 lines (26 ms instead of ~8 at T = 4801)
__device__ __forceinline__ void cp16_ca(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16z(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// fp32 += one fp16 (a single FHADD on sm_100; the half comes from either half of a 32-bit register)
__device__ __forceinline__ float fhadd(unsigned short h, float s) {
  float d;
  asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(s));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace ptx

#ifdef T6_CTATIME
__device__ long long t6_cta[65536][3];       // per CTA (debug builds only): clock64 at entry, clock64 at exit, SM id
#endif
#ifdef T6_TRACE
__device__ long long t6_trace[6][16][8];     // [role][tile][event] clock64 stamps of one CTA (debug builds only): roles 0..2 = softmax warp 0 of group g, 3 = MMA 1 issuer, 4 = MMA 3 issuer, 5 = loader
#define T6_STAMP(role, t, ev) do { if (blockIdx.x == T6_TRACE && (t) < 16 && lane == 0) t6_trace[role][t][ev] = clock64(); } while (0)
#else
#define T6_STAMP(role, t, ev) do { } while (0)
#endif
// +1 / -1: every (query, key) offset of tile t of the 64-query block at i0 lies at or beyond +512 / -512 (conformer.py:108 clamps
// the distance, so the rel-pos logit is a per-query constant); 0: the tile needs the R GEMM
__device__ __forceinline__ int t5_far(int i0, int t) {
  const int j0 = t * T5_KT;
  return (i0 - (j0 + T5_KT - 1) >= T5_MAXPOS) ? 1 : ((i0 + T5_BQ - 1 - j0 <= -T5_MAXPOS) ? -1 : 0);
}
__device__ __forceinline__ long long t5_seq_base(const SebSeq& sq, int seq) {
  return (long long)(seq / sq.inner) * sq.outer_stride + (seq % sq.inner);
}

// ------------------------------------------------------------------------------------------------------------------------
// Three query groups per CTA (one CTA per SM, all 512 TMEM columns).  The two-CTA kernel above is bound by the latency chain
// of its softmax warps (wait S -> tcgen05.ld -> STS / LDS -> FHADD -> max -> ex2 -> tcgen05.st -> arrive): two chains per
// scheduler leave every pipe under 45 %.  TMEM holds no third copy of (S | R | P | O) = 256 columns -- but R is live only from
// its MMA to the threads' tcgen05.ld (~300 of a tile's ~2500 clocks), so THREE groups of 128 accumulator rows (three consecutive
// 64-query blocks of one (sequence, head pair): they share every K / V tile and one 256-row E window) take turns on ONE R buffer:
//   TMEM: group g: S 128g | P 128g + 64 | O 128g + 96;  R 384..511 (shared, handed on through the R-free barrier)
// The strict (tile, group) order of the R users also staggers the three groups by a third of a tile period, so their LSU-,
// MUFU- and TMEM-bound phases interleave instead of colliding.
// Fringe key (kfr = 1 when n = 64 k + 1, the model's frame counts): the last key does not get a tile of its own; every thread computes its score on
// the FMA pipe before the loop (q . k + q . e from rows loaded ahead of the set-up barrier) and starts the online softmax from it: m = score, l = 1,
// O = v through tcgen05.st, first P . V with accumulate.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int T6_G = 3, T6_BQ = T6_G * T5_BQ;
constexpr int T6_W_LOAD = 4 * T6_G, T6_W_MMA1 = T6_W_LOAD + 1, T6_W_MMA3 = T6_W_LOAD + 2, T6_THREADS = (T6_W_MMA3 + 1) * 32;     // 480
constexpr int T6_STAGES = 4;
constexpr int T6_AEXP = 0, T6_QPL = T6_G * 8192, T6_STAGE0 = T6_QPL + T6_G * 4096;
constexpr int T6_KS = 0, T6_VS = 4096, T6_ES = 8192, T6_STAGE = 16384;           // E window: 256 rows (64 keys + 192 queries - 1 offsets)
constexpr int T6_RSCR = T6_STAGE0 + T6_STAGES * T6_STAGE;
constexpr int T6_SMEM = T6_RSCR + T6_G * 128 * T5_RPITCH + 128;
constexpr uint32_t T6_TR = 384;


template <int WM>
__global__ void __launch_bounds__(T6_THREADS, 1)
attention_tc3_kernel(const __half* __restrict__ qkvh, const __half* __restrict__ Eh, const SebSeq sq, int nqb, int nq, int ntiles, int kfr,
                     float* __restrict__ out) {
  extern __shared__ uint8_t t5_smraw[];
  __shared__ uint64_t bar_S[T6_G], bar_F[T6_G], bar_P[T6_G], bar_O[T6_G], bar_RF, full_bar[T6_STAGES], empty_bar[T6_STAGES];
  __shared__ uint32_t tmem_base_s;
#ifdef T6_CTATIME
  const long long cta_t0 = clock64();
#endif
  const uint32_t sm0 = (ptx::smem_u32(t5_smraw) + 127u) & ~127u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x % nqb, sh = blockIdx.x / nqb;
  const int hp = sh & 1, seq = sh >> 1;
  const int n = sq.n, i0 = qb * T6_BQ;
  const long long base = t5_seq_base(sq, seq);
  const int nt = n - kfr;                                               // keys covered by the tiles; keys nt .. n - 1 are the fringe (see attention_tc_launch)
  const int ng = min(T6_G, (nq - i0 + T5_BQ - 1) / T5_BQ);              // live groups of this CTA
  const __half* seq0 = qkvh + base * T5_ROWH;

  if (tid == 0) {
    for (int g = 0; g < T6_G; ++g) {
      ptx::mbar_init(&bar_S[g], 1); ptx::mbar_init(&bar_F[g], 128); ptx::mbar_init(&bar_P[g], 128); ptx::mbar_init(&bar_O[g], 1);
    }
    ptx::mbar_init(&bar_RF, 128);       // the 128 threads of the group that owns R have it in registers
    for (int s = 0; s < T6_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 32); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == T6_W_LOAD) ptx::tmem_alloc(&tmem_base_s, 512);

  // ---- loader state: K, V (8 + 8 chunks of 16 bytes per lane) and the 256-row E window (16 chunks per lane) of a tile
  const long long key_stride_h = sq.pos_stride * T5_ROWH;
  const __half* kv_lane = seq0 + (long long)(lane >> 2) * key_stride_h + 64 + hp * 32 + (lane & 3) * 8;
  const uint32_t k_dst = (uint32_t)(T6_KS + (lane & 3) * 128 + (lane >> 2) * 16);
  const uint32_t v_dst = (uint32_t)(T6_VS + (lane & 3) * 1024 + (lane >> 2) * 16);
  const uint32_t e_dst = (uint32_t)(T6_ES + (lane >> 4) * 256 + (lane & 1) * 128 + ((lane >> 1) & 7) * 16);
  const __half* e_lane = Eh + T5_MAXPOS * T5_D + (lane & 1) * 8;
  auto any_near = [&](int t) {            // does any live group need the R GEMM for tile t?
    bool r = false;
    for (int g = 0; g < ng; ++g) r = r || (t5_far(i0 + g * T5_BQ, t) == 0);
    return r;
  };
  auto issue_tile = [&](int t) {          // one commit group per call (empty past the last tile)
    if (t < ntiles) {
      const uint32_t st = sm0 + T6_STAGE0 + (uint32_t)((t % T6_STAGES) * T6_STAGE);
      const int j0 = t * T5_KT;
      const int key0 = j0 + (lane >> 2);
      const __half* src0 = kv_lane + (long long)j0 * key_stride_h;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool ok = key0 + 8 * k < nt;
        const __half* src = ok ? src0 + (long long)(8 * k) * key_stride_h : seq0;
        ptx::cp16z(st + k_dst + (uint32_t)(k * 512), src, ok ? 16u : 0u);
        ptx::cp16z(st + v_dst + (uint32_t)(k * 128), ok ? src + 64 : seq0, ok ? 16u : 0u);
      }
      if (any_near(t)) {                  // row c of the window = E[clamp(i0 + 191 - j0 - c)]; group g reads rows 64 (2 - g) .. + 127
        const int d0 = i0 + T6_BQ - 1 - j0 - (lane >> 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          int d = d0 - 16 * k;
          d = d < -T5_MAXPOS ? -T5_MAXPOS : (d > T5_MAXPOS ? T5_MAXPOS : d);
          ptx::cp16_ca(st + e_dst + (uint32_t)(k * 512), e_lane + d * T5_D);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (warp == T6_W_LOAD) {
    issue_tile(0);
    issue_tile(1);
  }

  const int gi = warp >> 2, q4 = warp & 3;                 // softmax warps: query group, TMEM lane quadrant
  const int i0g = i0 + gi * T5_BQ;
  uint4 qlo = make_uint4(0u, 0u, 0u, 0u), qhi = qlo;
  uint4 fk0 = qlo, fk1 = qlo, fv0 = qlo, fv1 = qlo, fe0 = qlo, fe1 = qlo;      // fringe key: its k / v rows and this row's table row, in flight across the set-up barrier
  if (warp < T6_W_LOAD && gi < ng) {
    // ---- operand rows of this thread: Aexp (q in its head's slot) and Q (k order permuted like the fragment-ordered E table)
    const int hl = q4 >> 1, par = q4 & 1, r = tid & 127;
    const int i = i0g + 2 * lane + par;
    if (i < nq) {
      const uint4* qp = reinterpret_cast<const uint4*>(seq0 + (long long)i * key_stride_h + (2 * hp + hl) * T5_D);
      qlo = __ldg(qp); qhi = __ldg(qp + 1);
    }
    if (kfr > 0) {
      const uint4* kp = reinterpret_cast<const uint4*>(seq0 + (long long)nt * key_stride_h + 64 + (2 * hp + hl) * T5_D);
      fk0 = __ldg(kp); fk1 = __ldg(kp + 1); fv0 = __ldg(kp + 8); fv1 = __ldg(kp + 9);         // v row = k row + 64 halfs
      int d = i - nt;
      d = d < -T5_MAXPOS ? -T5_MAXPOS : (d > T5_MAXPOS ? T5_MAXPOS : d);
      const uint4* ep = reinterpret_cast<const uint4*>(Eh + (d + T5_MAXPOS) * T5_D);
      fe0 = __ldg(ep); fe1 = __ldg(ep + 1);
    }
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t arow = sm0 + T6_AEXP + (uint32_t)(gi * 8192 + (r >> 3) * 512 + (r & 7) * 16);
    auto sts128 = [](uint32_t addr, uint4 v) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    };
    sts128(arow + 0 * 128, hl == 0 ? qlo : zero);
    sts128(arow + 1 * 128, hl == 0 ? qhi : zero);
    sts128(arow + 2 * 128, hl == 1 ? qlo : zero);
    sts128(arow + 3 * 128, hl == 1 ? qhi : zero);
    const uint32_t qrow = sm0 + T6_QPL + (uint32_t)(gi * 4096 + (r >> 3) * 256 + (r & 7) * 16);
    sts128(qrow, make_uint4(qlo.x, qhi.x, qlo.y, qhi.y));
    sts128(qrow + 128, make_uint4(qlo.z, qhi.z, qlo.w, qhi.w));
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == T6_W_LOAD) {
    // ================= loader warp (continued) =================
    for (int t = 2; t < ntiles + 2; ++t) {
      {
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // this lane's chunks of tile t - 2 have landed
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&full_bar[(t - 2) % T6_STAGES]);
      }
      if (t >= T6_STAGES && t < ntiles) {
        if (lane == 0) ptx::mbar_wait_lean<WM>(&empty_bar[t % T6_STAGES], (uint32_t)(t / T6_STAGES - 1) & 1u);
        __syncwarp();
      }
      issue_tile(t);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == T6_W_MMA1) {
    // ================= MMA 1 issuer: strict (tile, group) order; the R GEMM of a group waits until the previous owner of R has read it =================
    if (lane == 0) {
      constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t IDESC_R = ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int ruse = 0;                        // R GEMMs issued so far = phases of bar_RF
      for (int t = 0; t < ntiles; ++t) {
        const int slot = t % T6_STAGES;
        const uint32_t st = sm0 + T6_STAGE0 + (uint32_t)(slot * T6_STAGE);
        const uint64_t bk = ptx::umma_desc_ns(st + T6_KS, 128, 512);
        ptx::mbar_wait_lean<WM>(&full_bar[slot], (uint32_t)(t / T6_STAGES) & 1u);
        T6_STAMP(3, t, 6);
        for (int g = 0; g < ng; ++g) {
          const uint32_t tS = tmem_base + (uint32_t)(128 * g);
          const uint64_t a0 = ptx::umma_desc_ns(sm0 + T6_AEXP + g * 8192, 128, 512);
          if (t > 0) ptx::mbar_wait_lean<WM>(&bar_F[g], (uint32_t)(t - 1) & 1u);
          const bool near = t5_far(i0 + g * T5_BQ, t) == 0;
          if (near && ruse > 0) ptx::mbar_wait_lean<WM>(&bar_RF, (uint32_t)(ruse - 1) & 1u);
          ptx::tc_fence_after();
          T6_STAMP(3, t, 2 * g);
          ptx::mma_f16_ss(tS, a0, bk, IDESC_S, 0u);
          ptx::mma_f16_ss(tS, a0 + (256 >> 4), bk + (256 >> 4), IDESC_S, 1u);
          if (near) {
            ptx::mma_f16_ss(tmem_base + T6_TR, ptx::umma_desc_ns(sm0 + T6_QPL + g * 4096, 128, 256),
                            ptx::umma_desc_ns(st + T6_ES + (T6_G - 1 - g) * 2048, 128, 256), IDESC_R, 0u);
            ++ruse;
          }
          ptx::tc_commit(&bar_S[g]);
          T6_STAMP(3, t, 2 * g + 1);
        }
      }
    }
  } else if (warp == T6_W_MMA3) {
    // ================= MMA 3 issuer: O_g += P_g(t) . V(t) in (tile, group) order; the last group's commit frees the ring slot =================
    if (lane == 0) {
      constexpr uint32_t IDESC_O = (1u << 4) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int t = 0; t < ntiles; ++t) {
        const int slot = t % T6_STAGES;
        const uint64_t bv = ptx::umma_desc_ns(sm0 + T6_STAGE0 + (uint32_t)(slot * T6_STAGE) + T6_VS, 128, 1024);
        const int nks = (nt - t * T5_KT <= 16) ? 1 : 4;
        const uint32_t acc0 = (t > 0 || kfr > 0) ? 1u : 0u;          // with fringe keys the accumulator starts from their contribution
        for (int g = 0; g < ng; ++g) {
          const uint32_t tP = tmem_base + (uint32_t)(128 * g + 64), tO = tmem_base + (uint32_t)(128 * g + 96);
          ptx::mbar_wait_lean<WM>(&bar_P[g], (uint32_t)t & 1u);
          ptx::tc_fence_after();
          T6_STAMP(4, t, 2 * g);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < nks) ptx::mma_f16_ts(tO, tP + (uint32_t)(ks * 8), bv + (uint64_t)((ks * 256) >> 4), IDESC_O, ks ? 1u : acc0);
          ptx::tc_commit(&bar_O[g]);
          T6_STAMP(4, t, 2 * g + 1);
        }
        ptx::tc_commit(&empty_bar[slot]);
      }
    }
  } else if (gi < ng) {
    // ================= softmax threads: one accumulator row each =================
    const int hl = q4 >> 1, par = q4 & 1;
    const int i = i0g + 2 * lane + par;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + (uint32_t)(128 * gi), tP = tS + 64u, tO = tS + 96u, tR = tmem_base + lane_sel + T6_TR;
    const uint32_t myrow = sm0 + T6_RSCR + (uint32_t)(tid * T5_RPITCH);
    const uint32_t rd = myrow + (uint32_t)((31 - lane) * 4);
    uint64_t* const bS = &bar_S[gi]; uint64_t* const bF = &bar_F[gi]; uint64_t* const bP = &bar_P[gi]; uint64_t* const bO = &bar_O[gi];
    float m = 0.f, l = 0.f;
    float c_hi = 0.f, c_lo = 0.f;
    {
      const uint32_t qperm[8] = {qlo.x, qhi.x, qlo.y, qhi.y, qlo.z, qhi.z, qlo.w, qhi.w};      // k order of the fragment-ordered table rows
      auto dot_e = [&](const __half* erow) {                  // q . (one row of the packed table), fp32
        const uint4 e0 = __ldg(reinterpret_cast<const uint4*>(erow)), e1 = __ldg(reinterpret_cast<const uint4*>(erow) + 1);
        const uint32_t ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qperm[k])), ef = __half22float2(*reinterpret_cast<const __half2*>(&ee[k]));
          acc = fmaf(qf.x, ef.x, fmaf(qf.y, ef.y, acc));
        }
        return acc;
      };
      if (n > T5_MAXPOS) {
        c_hi = dot_e(Eh + 2 * T5_MAXPOS * T5_D);
        c_lo = dot_e(Eh);
      }
      if (kfr > 0) {
        // ---- fringe key nt = n - 1: the online softmax starts from it (m, l and the O accumulator), computed on the FMA pipe from the rows
        // loaded before the set-up barrier
        const uint32_t qnat[8] = {qlo.x, qlo.y, qlo.z, qlo.w, qhi.x, qhi.y, qhi.z, qhi.w};
        const uint32_t kk[8] = {fk0.x, fk0.y, fk0.z, fk0.w, fk1.x, fk1.y, fk1.z, fk1.w}, vv[8] = {fv0.x, fv0.y, fv0.z, fv0.w, fv1.x, fv1.y, fv1.z, fv1.w};
        const uint32_t ee[8] = {fe0.x, fe0.y, fe0.z, fe0.w, fe1.x, fe1.y, fe1.z, fe1.w};
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qnat[k])), kf = __half22float2(*reinterpret_cast<const __half2*>(&kk[k]));
          const float2 qp2 = __half22float2(*reinterpret_cast<const __half2*>(&qperm[k])), ef = __half22float2(*reinterpret_cast<const __half2*>(&ee[k]));
          acc = fmaf(qf.x, kf.x, fmaf(qf.y, kf.y, fmaf(qp2.x, ef.x, fmaf(qp2.y, ef.y, acc))));
        }
        m = acc;                             // the single fringe score is its own maximum: p = 1
        l = 1.f;
        uint32_t ow[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&vv[k]));
          ow[2 * k] = __float_as_uint(vf.x); ow[2 * k + 1] = __float_as_uint(vf.y);
        }
        ptx::tmem_st16(tO + (uint32_t)(hl * 16), ow);
        ptx::tmem_st_wait5();
      }
    }
    auto tile_body = [&](auto nk_tag, auto far_tag, int t, float cadd) {
      constexpr int NK = decltype(nk_tag)::value;
      constexpr bool FAR = decltype(far_tag)::value;
      constexpr int NWR = NK == 64 ? 64 : 40;
      if (q4 == 0) T6_STAMP(gi, t, 0);
      ptx::mbar_wait_lean<WM>(bS, (uint32_t)t & 1u);
      ptx::tc_fence_after();
      if (q4 == 0) T6_STAMP(gi, t, 1);
      if (!FAR) {
        uint32_t w[64];
        ptx::tmem_ld32_pack16<0>(tR, w);
        if (NK == 64) ptx::tmem_ld32_pack16<32>(tR + 64u, w); else ptx::tmem_ld16_pack16(tR + 64u, w + 32);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bar_RF);       // R may go to the next group
        if (q4 == 0) T6_STAMP(gi, t, 2);
#pragma unroll
        for (int q = 0; q < NWR / 4; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(myrow + q * 16), "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]) : "memory");
      }
      uint32_t sb[NK];
      if (NK == 64) { ptx::tmem_ld32<0>(tS, sb); ptx::tmem_ld32<32>(tS + 32u, sb); } else ptx::tmem_ld16(tS, sb);
      uint32_t x[NK / 2 + 1];
      if (!FAR) {
#pragma unroll
        for (int k = 0; k < NK / 2 + 1; ++k) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x[k]) : "r"(rd + k * 4) : "memory");
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bF);              // S sits in registers: the next tile's MMA 1 may overwrite it
      if (q4 == 0) T6_STAMP(gi, t, 3);
      float s[NK];
      if (FAR) {
#pragma unroll
        for (int jj = 0; jj < NK; ++jj) s[jj] = __uint_as_float(sb[jj]);
      } else if (par) {
#pragma unroll
        for (int p = 0; p < NK / 2; ++p) {
          s[2 * p] = ptx::fhadd((unsigned short)(x[p] & 0xffffu), __uint_as_float(sb[2 * p]));
          s[2 * p + 1] = ptx::fhadd((unsigned short)(x[p] >> 16), __uint_as_float(sb[2 * p + 1]));
        }
      } else {
#pragma unroll
        for (int p = 0; p < NK / 2; ++p) {
          s[2 * p] = ptx::fhadd((unsigned short)(x[p] >> 16), __uint_as_float(sb[2 * p]));
          s[2 * p + 1] = ptx::fhadd((unsigned short)(x[p + 1] & 0xffffu), __uint_as_float(sb[2 * p + 1]));
        }
      }
      const int rem = nt - t * T5_KT;
      if (rem < NK) {
#pragma unroll
        for (int jj = 0; jj < NK; ++jj)
          if (jj >= rem) s[jj] = -1e30f;
      }
      float mx;
      {
        constexpr int N1 = (NK + 2) / 3;
        float a[N1];
#pragma unroll
        for (int k = 0; k < NK / 3; ++k) a[k] = ptx::fmax3(s[3 * k], s[3 * k + 1], s[3 * k + 2]);
        a[N1 - 1] = s[NK - 1];
        mx = a[0];
#pragma unroll
        for (int k = 1; k + 1 < N1; k += 2) mx = ptx::fmax3(mx, a[k], a[k + 1]);
        if ((N1 & 1) == 0) mx = fmaxf(mx, a[N1 - 1]);
        if (FAR) mx += cadd;
      }
      if (q4 == 0) T6_STAMP(gi, t, 4);
      if (t > 0) {
        ptx::mbar_wait_lean<WM>(bO, (uint32_t)(t - 1) & 1u);
        ptx::tc_fence_after();
      }
      if (q4 == 0) T6_STAMP(gi, t, 5);
      if (t == 0 && kfr == 0) {
        m = mx;
      } else {
        const bool need = mx > m + T5_LAZY;
        if (__any_sync(0xffffffffu, need)) {
          const float mn = need ? mx : m;
          const float corr = ptx::ex2f(m - mn);
          m = mn;
          l *= corr;
          uint32_t o[32];
          ptx::tmem_ld32<0>(tO, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * corr);
          ptx::tmem_st32(tO, o);
        }
      }
      uint32_t pw[NK / 2];
      float2 la = make_float2(0.f, 0.f), lb = la;
      const float2 negm = FAR ? make_float2(cadd - m, cadd - m) : make_float2(-m, -m);
#pragma unroll
      for (int p = 0; p < NK / 2; p += 2) {
        const float2 a = __fadd2_rn(make_float2(s[2 * p], s[2 * p + 1]), negm), b = __fadd2_rn(make_float2(s[2 * p + 2], s[2 * p + 3]), negm);
        const float2 pa = make_float2(ptx::ex2f(a.x), ptx::ex2f(a.y)), pb = make_float2(ptx::ex2f(b.x), ptx::ex2f(b.y));
        la = __fadd2_rn(la, pa); lb = __fadd2_rn(lb, pb);
        const __half2 h0 = __floats2half2_rn(pa.x, pa.y), h1 = __floats2half2_rn(pb.x, pb.y);
        pw[p] = *reinterpret_cast<const uint32_t*>(&h0);
        pw[p + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
      l += (la.x + la.y) + (lb.x + lb.y);
      if (q4 == 0) T6_STAMP(gi, t, 6);
      if (NK == 64) ptx::tmem_st32(tP, pw); else ptx::tmem_st8u(tP, pw);
      ptx::tmem_st_wait5();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bP);
      if (q4 == 0) T6_STAMP(gi, t, 7);
    };
    if (i0g + par >= nq) {               // every query row of this warp lies past the sequence: keep the protocol, skip the work
      for (int t = 0; t < ntiles; ++t) {
        ptx::mbar_wait_lean<WM>(bS, (uint32_t)t & 1u);
        if (t5_far(i0g, t) == 0) ptx::mbar_arrive(&bar_RF);
        ptx::mbar_arrive(bF);
        if (t > 0) ptx::mbar_wait_lean<WM>(bO, (uint32_t)(t - 1) & 1u);
        ptx::mbar_arrive(bP);
      }
    } else {
      for (int t = 0; t < ntiles; ++t) {
        const int far = t5_far(i0g, t);
        const bool tail = nt - t * T5_KT <= 16;
        if (far && tail) tile_body(std::integral_constant<int, 16>{}, std::true_type{}, t, far > 0 ? c_hi : c_lo);
        else if (far) tile_body(std::integral_constant<int, 64>{}, std::true_type{}, t, far > 0 ? c_hi : c_lo);
        else if (tail) tile_body(std::integral_constant<int, 16>{}, std::false_type{}, t, 0.f);
        else tile_body(std::integral_constant<int, 64>{}, std::false_type{}, t, 0.f);
      }
    }
    ptx::mbar_wait_lean<WM>(bO, (uint32_t)(ntiles - 1) & 1u);
    ptx::tc_fence_after();
    uint32_t o[16];
    ptx::tmem_ld16(tO + (uint32_t)(hl * 16), o);
    ptx::tmem_ld_wait();
    if (i < nq) {
      const float inv = 1.0f / l;
      float* op = out + (base + (long long)i * sq.pos_stride) * 64 + (2 * hp + hl) * T5_D;
#pragma unroll
      for (int c = 0; c < 16; c += 4)
        *reinterpret_cast<float4*>(op + c) = make_float4(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv,
                                                         __uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == T6_W_LOAD) ptx::tmem_dealloc(tmem_base, 512);
#ifdef T6_CTATIME
  if (threadIdx.x == 0 && blockIdx.x < 65536) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    t6_cta[blockIdx.x][0] = cta_t0; t6_cta[blockIdx.x][1] = clock64(); t6_cta[blockIdx.x][2] = smid;
  }
#endif
}


#ifdef T6_CTATIME
extern "C" int seb200_t6_cta(long long* host) { return (int)cudaMemcpyFromSymbol(host, t6_cta, sizeof(t6_cta)); }
#endif
#ifdef T6_TRACE
extern "C" int seb200_t6_trace(long long* host) { return (int)cudaMemcpyFromSymbol(host, t6_trace, sizeof(t6_trace)); }
#endif

int attention_tc_launch(const __half* qkvh, const __half* Eh, const SebSeq* seq, float* out, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, T6_SMEM);
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  const int n = seq->n;
  // The model's frame counts are 64 k + 1 (T = L / 100 + 1): instead of a last key tile holding ONE key, that key enters the threads' initial
  // softmax state (m, l, O) on the FMA pipe -- one synchronisation round less per CTA (-2 % at n = 641, profiles/r2/attention_latency_analysis.txt)
  const int kfr = (n > T5_KT && n % T5_KT == 1) ? 1 : 0;
  const int ntiles = (n - kfr + T5_KT - 1) / T5_KT;
  const int nqb = (n + T6_BQ - 1) / T6_BQ;
  const long long nb = (long long)seq->nseq * 2 * nqb;
  SEB_REQUIRE(nb < 2147483647LL, SEB_EINVAL, "attention: grid too large");
  attention_tc3_kernel<0><<<(unsigned)nb, T6_THREADS, T6_SMEM, st>>>(qkvh, Eh, *seq, nqb, n, ntiles, kfr, out);
  SEB_CHECK_LAUNCH("attention_tc3_kernel");
  return 0;
}

}  // namespace seb
