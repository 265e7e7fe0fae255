// ffn_fused.cu -- one kernel for the conformer feed-forward half-step (conformer.py:128-145 wrapped by PreNorm and
// Scale(0.5), conformer.py:53-71, 201-202, 207, 210), optionally followed by post_norm and the TSCB outer residual
// (conformer.py:211, generator.py:70,72):
//
//     y   = x + alpha * ( W2 . swish( W1 . LayerNorm(x) + b1 ) + b2 )
//     out = post ? LayerNorm_post(y) + resid2 : y
//
// v3.  ncu of v2 showed the kernel bound by SHARED-MEMORY bandwidth and by all 16 epilogue warps marching in lock step
// (MUFU pipe 32 % busy): with the 3-product bf16 split every tcgen05.mma re-read its 128 x 16 A slice from shared
// memory (N = 64 is too narrow to amortise it), the hidden tile H made a round trip through shared memory, and the
// loaders waited on eight serialised global loads per tile.  v3 keeps BOTH A operands in tensor memory (the
// `[a_tmem]` form of tcgen05.mma), so shared memory only serves the resident weights:
//
//   copy warp      : cp.async 16 B of the next tile's raw fp32 rows into a 2-slot swizzled staging tile; loads W1|W2 once
//   4 LN warps     : thread = row.  staged row -> LayerNorm -> bf16 hi|lo -> tcgen05.st into XA[s]            (TMEM)
//   MMA1 issuer    : ACC1[g] = XA[s] . W1[64q : 64q+64]^T      (A from TMEM, W1 resident in smem; 3 products)
//   16 mid warps   : two groups of 8 that alternate hidden chunks (g = q & 1), so one group's MUFU phase overlaps the
//                    other's TMEM traffic:  tcgen05.ld ACC1[g] -> +b1 -> swish -> bf16 hi|lo -> tcgen05.st H[g]  (TMEM)
//   MMA2 issuer    : ACC2[ab] += H[g] . W2[:, 64q : 64q+64]^T  (A from TMEM, W2 resident in smem)
//   4 final warps  : thread = row.  tcgen05.ld ACC2 -> *alpha + b2 + x (staged row) -> row statistics -> staged in place ->
//                    coalesced copy-out (normalise + resid2 when post-norm is requested)
//
// TMEM map (512 columns): XA[2] | ACC1[2] | H[2] | ACC2[2], 64 columns each.  All hand-offs are mbarriers; the two MMA
// issuers are separate threads so that GEMM 1 of chunk q + 2 never queues behind GEMM 2 of chunk q.
#include "gemm_engine.cuh"

namespace seb {

struct FfnArgs {
  const float* x; float* out; int M;
  const float* ln_g; const float* ln_b;
  const uint8_t* w1; const float* b1;
  const uint8_t* w2; const float* b2;
  float alpha;
  const float* pn_g; const float* pn_b; const float* resid2;
};

constexpr int F3_LN_WARPS = 4, F3_MID_WARPS = 16, F3_FIN_WARPS = 4;
constexpr int F3_W_MID0 = F3_LN_WARPS;                       // first mid warp (multiple of 4: warp % 4 = TMEM lane quarter)
constexpr int F3_W_FIN0 = F3_W_MID0 + F3_MID_WARPS;          // 20
constexpr int F3_W_MMA1 = F3_W_FIN0 + F3_FIN_WARPS;          // 24
constexpr int F3_W_MMA2 = F3_W_MMA1 + 1;                     // 25
constexpr int F3_W_COPY = F3_W_MMA2 + 1;                     // 26
constexpr int F3_THREADS = (F3_W_COPY + 1) * 32;             // 864
constexpr int F3_WBLK = 2 * 64 * 128;                        // 16 KB: hi|lo image of a 64-row x 64-k weight block
constexpr int F3_XSLOT = BM * 256;                           // 32 KB: raw fp32 rows of one tile (16-byte chunks XOR-swizzled by row & 7)
constexpr int F3_NSLOT = 3;                                  // staging slots: a slot lives from the copy to the end of its tile, so the slot
                                                             // count caps the tiles in flight (two slots left the DRAM latency in the chain)
constexpr int F3_TAB_FLOATS = 256 + 64;                      // b1 | b2  (LayerNorm parameters come through L1, row statistics by warp shuffle)
constexpr int F3_SMEM = 1024 + 8 * F3_WBLK + F3_NSLOT * F3_XSLOT + F3_TAB_FLOATS * 4;
static_assert(F3_SMEM + 256 <= 232448, "ffn_fused: shared memory over the 227 KB per-CTA limit (dynamic + the static mbarriers)");
// TMEM columns
constexpr uint32_t F3_XA = 0, F3_ACC1 = 128, F3_H = 256, F3_ACC2 = 384;

namespace ptx {
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace ptx

__device__ __forceinline__ float4 lds4(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }

// swish of 16 accumulator columns (+ bias) -> 8 packed bf16 hi columns + 8 packed lo columns.
// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2, sm_100): the bias add, the exponent scaling, 1 + e, v * sigmoid and the
// hi/lo residual each cost one instruction per PAIR of hidden elements -- 6.5 instead of ~9.5 instructions per element in the
// loop that bounds the kernel (issue slots); the two MUFU ops per element stay scalar.
#ifndef SEB_FFN_EXP
#define SEB_FFN_EXP 0
#endif
#ifndef SEB_FFN_PACKED
#define SEB_FFN_PACKED 1
#endif
__device__ __forceinline__ void swish_split_pair(float a0, float a1, float2 b, uint32_t& hi, uint32_t& lo) {
  const float2 v = __fadd2_rn(make_float2(a0, a1), b);
  const float2 t = __fmul2_rn(v, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t.y));
  const float2 d = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
#if SEB_FFN_EXP == 1          // timing experiment only: no MUFU work in the mid epilogue
  const float2 s = v;
#else
  const float2 s = __fmul2_rn(v, make_float2(r0, r1));
#endif
  __nv_bfloat162 h = __floats2bfloat162_rn(s.x, s.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float2 hf = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  const float2 l = __ffma2_rn(hf, make_float2(-1.0f, -1.0f), s);
  __nv_bfloat162 lb = __floats2bfloat162_rn(l.x, l.y);
  lo = *reinterpret_cast<uint32_t*>(&lb);
}

__device__ __forceinline__ void swish_split16(const uint32_t (&r)[16], const float* bias, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * j);
#if SEB_FFN_PACKED
    swish_split_pair(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), make_float2(b.x, b.y), hi[2 * j], lo[2 * j]);
    swish_split_pair(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]), make_float2(b.z, b.w), hi[2 * j + 1], lo[2 * j + 1]);
#else
    float v0 = __uint_as_float(r[4 * j]) + b.x, v1 = __uint_as_float(r[4 * j + 1]) + b.y;
    float v2 = __uint_as_float(r[4 * j + 2]) + b.z, v3 = __uint_as_float(r[4 * j + 3]) + b.w;
    v0 *= sigmoidf_acc(v0); v1 *= sigmoidf_acc(v1); v2 *= sigmoidf_acc(v2); v3 *= sigmoidf_acc(v3);
    split_bf16x2(v0, v1, hi[2 * j], lo[2 * j]);
    split_bf16x2(v2, v3, hi[2 * j + 1], lo[2 * j + 1]);
#endif
  }
}

__global__ void __launch_bounds__(F3_THREADS, 1) ffn_fused_kernel(const FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full, x_full[F3_NSLOT], x_empty[F3_NSLOT], xa_full[2], xa_empty[2], acc1_full[2], acc1_empty[2], h_full[2], h_empty[2],
      acc2_full[2], acc2_empty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space
  uint8_t* sW1 = smem;                          // 4 blocks x (hi | lo)   64 KB  resident
  uint8_t* sW2 = sW1 + 4 * F3_WBLK;             // 4 blocks x (hi | lo)   64 KB  resident
  uint8_t* sX = sW2 + 4 * F3_WBLK;              // [2] raw fp32 tiles     64 KB
  float* sTab = reinterpret_cast<float*>(sX + F3_NSLOT * F3_XSLOT);
  float* sB1 = sTab; float* sB2 = sTab + 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (a.M + BM - 1) / BM;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const bool post = a.pn_g != nullptr;

  if (tid == 0) {
    ptx::mbar_init(&w_full, 1);
    for (int i = 0; i < F3_NSLOT; ++i) { ptx::mbar_init(&x_full[i], 32); ptx::mbar_init(&x_empty[i], F3_FIN_WARPS * 32); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&xa_full[i], F3_LN_WARPS * 32);          ptx::mbar_init(&xa_empty[i], 1);
      ptx::mbar_init(&acc1_full[i], 1);                       ptx::mbar_init(&acc1_empty[i], (F3_MID_WARPS / 2) * 32);
      ptx::mbar_init(&h_full[i], (F3_MID_WARPS / 2) * 32);    ptx::mbar_init(&h_empty[i], 1);
      ptx::mbar_init(&acc2_full[i], 1);                       ptx::mbar_init(&acc2_empty[i], F3_FIN_WARPS * 32);
    }
    ptx::fence_barrier_init();
  }
  for (int i = tid; i < F3_TAB_FLOATS; i += F3_THREADS) sTab[i] = (i < 256) ? a.b1[i] : a.b2[i - 256];
  if (warp == F3_W_MMA1) ptx::tmem_alloc(&tmem_base_s, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < F3_LN_WARPS) {
    // ================= LayerNorm warps: thread = row; staged x row -> LN -> bf16 hi|lo -> XA[s] in TMEM =================
    const int row = tid, sw = row & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1, xsl = it % F3_NSLOT;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      const uint8_t* xr = sX + xsl * F3_XSLOT + row * 256;
      ptx::mbar_wait(&x_full[xsl], (uint32_t)(it / F3_NSLOT) & 1u);
      // pass A: shifted one-pass statistics (shift = first element of the row; exact for constant rows)
      float sum = 0.f, sq = 0.f;
      const float x0 = lds4(xr + ((0 ^ sw) << 4)).x;
      {
        const float2 nx0 = make_float2(-x0, -x0);
        float2 s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, q2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int c = 0; c < 16; ++c) stats_acc4(lds4(xr + ((c ^ sw) << 4)), nx0, s2[c & 1], q2[c & 1]);
        sum = (s2[0].x + s2[0].y) + (s2[1].x + s2[1].y);
        sq = (q2[0].x + q2[0].y) + (q2[1].x + q2[1].y);
      }
      const float md = sum * (1.0f / 64.0f);
      const float var = fmaxf(sq * (1.0f / 64.0f) - md * md, 0.f);
      const float rstd = 1.0f / sqrtf(var + 1e-5f);
      const float mean = x0 + md;
      ptx::mbar_wait(&xa_empty[s], ph ^ 1u);
      ptx::tc_fence_after();
      const uint32_t xa = lane_base + F3_XA + (uint32_t)(s * 64);
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {          // 16 k-values -> 8 hi + 8 lo packed columns
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c16 * 4 + j;
          const float4 v = lds4(xr + ((c ^ sw) << 4));
          const float4 gg = ldg4(a.ln_g + c * 4), bb = ldg4(a.ln_b + c * 4);      // warp-uniform, L1-resident (no room left in shared memory)
          float2 y01, y23;
          ln_apply4(v, mean, rstd, gg, bb, y01, y23);
          split_bf16x2(y01.x, y01.y, hi[2 * j], lo[2 * j]);
          split_bf16x2(y23.x, y23.y, hi[2 * j + 1], lo[2 * j + 1]);
        }
        ptx::tmem_st8(xa + (uint32_t)(c16 * 8), hi);
        ptx::tmem_st8(xa + 32u + (uint32_t)(c16 * 8), lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&xa_full[s]);
    }
  } else if (warp < F3_W_FIN0) {
    // ================= mid-epilogue warps: ACC1[g] -> +b1 -> swish -> bf16 hi|lo -> H[g] (TMEM) =================
    const int e = warp - F3_W_MID0;
    const int wq = e & 3, g = (e >> 2) & 1, hf = e >> 3;          // lane quarter, chunk parity handled, column half
    const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
    const uint32_t t_acc = lane_base + F3_ACC1 + (uint32_t)(g * 64 + hf * 32);
    const uint32_t t_h = lane_base + F3_H + (uint32_t)(g * 64 + hf * 16);
    for (int it = 0; it < my_tiles; ++it) {
#pragma unroll 1
      for (int qq = 0; qq < 2; ++qq) {
        const int q = 2 * qq + g;
        const uint32_t u = (uint32_t)(2 * it + qq);               // use count of ACC1[g] / H[g]
        const float* bias = sB1 + q * 64 + hf * 32;
        ptx::mbar_wait(&acc1_full[g], u & 1u);
        ptx::tc_fence_after();
        uint32_t r0[16], r1[16];
        ptx::tmem_ld16_nowait(t_acc, r0);
        ptx::tmem_ld16_nowait(t_acc + 16u, r1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&acc1_empty[g]);
        uint32_t hi[16], lo[16];
        swish_split16(r0, bias, hi, lo);
        swish_split16(r1, bias + 16, hi + 8, lo + 8);
        ptx::mbar_wait(&h_empty[g], (u & 1u) ^ 1u);
        ptx::tc_fence_after();
        ptx::tmem_st8(t_h, hi);
        ptx::tmem_st8(t_h + 8u, hi + 8);
        ptx::tmem_st8(t_h + 32u, lo);
        ptx::tmem_st8(t_h + 40u, lo + 8);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&h_full[g]);
      }
    }
  } else if (warp < F3_W_MMA1) {
    // ================= final warps: thread = row; ACC2 -> *alpha + b2 + x -> (post-norm statistics) -> coalesced store =================
    const int fw = warp - F3_W_FIN0;
    const int row = fw * 32 + lane, sw = row & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(fw * 32) << 16);
    const int cc = lane & 15;                                     // 16-byte chunk handled in the copy-out
    const float4 pg = post ? ldg4(a.pn_g + cc * 4) : make_float4(1.f, 1.f, 1.f, 1.f), pb = post ? ldg4(a.pn_b + cc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
      const int xsl = it % F3_NSLOT, ab = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      uint8_t* xs = sX + xsl * F3_XSLOT;
      uint8_t* xr = xs + row * 256;
      // the residual rows of this tile are consumed by the copy-out below: start pulling them into L2 now
      if (post && a.resid2 != a.x) {
        const long long mrow = (long long)m0 + row;
        if (mrow < (long long)a.M) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.resid2 + mrow * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.resid2 + mrow * 64 + 32));
        }
      }
      ptx::mbar_wait(&acc2_full[ab], ph);
      ptx::tc_fence_after();
      const uint32_t t_acc = lane_base + F3_ACC2 + (uint32_t)(ab * 64);
      float y0s = 0.f, sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c16 = 0; c16 < (SEB_FFN_EXP == 4 ? 0 : 4); ++c16) {
        uint32_t r[16];
        ptx::tmem_ld16_nowait(t_acc + (uint32_t)(c16 * 16), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c16 * 4 + j;
          uint8_t* p = xr + ((c ^ sw) << 4);
          const float4 xv = lds4(p);
          const float4 b2 = *reinterpret_cast<const float4*>(sB2 + c * 4);
          float4 y;
          y.x = fmaf(a.alpha, __uint_as_float(r[4 * j]) + b2.x, xv.x);
          y.y = fmaf(a.alpha, __uint_as_float(r[4 * j + 1]) + b2.y, xv.y);
          y.z = fmaf(a.alpha, __uint_as_float(r[4 * j + 2]) + b2.z, xv.z);
          y.w = fmaf(a.alpha, __uint_as_float(r[4 * j + 3]) + b2.w, xv.w);
          if (c == 0) y0s = y.x;
          const float d0 = y.x - y0s, d1 = y.y - y0s, d2 = y.z - y0s, d3 = y.w - y0s;
          sum += (d0 + d1) + (d2 + d3);
          sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
          *reinterpret_cast<float4*>(p) = y;
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc2_empty[ab]);
      float my_mean, my_rstd;                  // this thread's row; the copy-out fetches a row's pair from its owner lane by shuffle
      {
        const float md = sum * (1.0f / 64.0f);
        const float var = fmaxf(sq * (1.0f / 64.0f) - md * md, 0.f);
        my_mean = y0s + md;
        my_rstd = 1.0f / sqrtf(var + 1e-5f);
      }
      __syncwarp();
      // coalesced copy-out of this warp's 32 rows: half a warp per row, one 16-byte chunk per lane.  Loads of a group of four
      // rows are issued before its stores: `out` may alias `resid2`, so the compiler cannot hoist them itself and the
      // residual reads would otherwise serialise into sixteen DRAM round trips per tile.
#pragma unroll 1
      for (int i0 = 0; i0 < (SEB_FFN_EXP == 2 ? 0 : 16); i0 += 4) {
        float4 yv[4], r2[4];
        float2 st[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = fw * 32 + 2 * (i0 + u) + (lane >> 4);
          const int m = m0 + rr;
          yv[u] = lds4(xs + rr * 256 + ((cc ^ (rr & 7)) << 4));
          st[u] = make_float2(__shfl_sync(0xffffffffu, my_mean, rr & 31), __shfl_sync(0xffffffffu, my_rstd, rr & 31));
          r2[u] = (post && m < a.M) ? *reinterpret_cast<const float4*>(a.resid2 + (long long)m * 64 + cc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = m0 + fw * 32 + 2 * (i0 + u) + (lane >> 4);
          float4 y = yv[u];
          if (post) {
            y.x = fmaf((y.x - st[u].x) * st[u].y, pg.x, pb.x) + r2[u].x; y.y = fmaf((y.y - st[u].x) * st[u].y, pg.y, pb.y) + r2[u].y;
            y.z = fmaf((y.z - st[u].x) * st[u].y, pg.z, pb.z) + r2[u].z; y.w = fmaf((y.w - st[u].x) * st[u].y, pg.w, pb.w) + r2[u].w;
          }
          if (m < a.M) st4(a.out + (long long)m * 64 + cc * 4, y);
        }
      }
      __syncwarp();
      ptx::mbar_arrive(&x_empty[xsl]);          // this thread's reads of the staging slot are done
    }
  } else if (warp == F3_W_MMA1) {
    // ================= GEMM 1 issuer: ACC1[g] = XA[s] . W1[q]^T =================
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t uW1 = ptx::smem_u32(sW1);
      ptx::mbar_wait(&w_full, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&xa_full[s], (uint32_t)(it >> 1) & 1u);
        for (int q = 0; q < 4; ++q) {
          const int g = q & 1;
          const uint32_t u = (uint32_t)(2 * it + (q >> 1));
          ptx::mbar_wait(&acc1_empty[g], (u & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + F3_ACC1 + (uint32_t)(g * 64);
          const uint32_t a_hi = tmem_base + F3_XA + (uint32_t)(s * 64), a_lo = a_hi + 32u;
          const uint64_t w_hi = ptx::umma_desc_sw128(uW1 + q * F3_WBLK), w_lo = ptx::umma_desc_sw128(uW1 + q * F3_WBLK + 64 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ptx::mma_bf16_ts(d, a_lo + (uint32_t)(k * 8), w_hi + ko, IDESC, k == 0 ? 0u : 1u);
            ptx::mma_bf16_ts(d, a_hi + (uint32_t)(k * 8), w_lo + ko, IDESC, 1u);
            ptx::mma_bf16_ts(d, a_hi + (uint32_t)(k * 8), w_hi + ko, IDESC, 1u);
          }
          ptx::tc_commit(&acc1_full[g]);
          if (q == 3) ptx::tc_commit(&xa_empty[s]);
        }
      }
    }
  } else if (warp == F3_W_MMA2) {
    // ================= GEMM 2 issuer: ACC2[ab] += H[g] . W2[:, q]^T =================
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t uW2 = ptx::smem_u32(sW2);
      ptx::mbar_wait(&w_full, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int ab = it & 1;
        for (int q = 0; q < 4; ++q) {
          const int g = q & 1;
          const uint32_t u = (uint32_t)(2 * it + (q >> 1));
          if (q == 0) ptx::mbar_wait(&acc2_empty[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);
          ptx::mbar_wait(&h_full[g], u & 1u);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + F3_ACC2 + (uint32_t)(ab * 64);
          const uint32_t a_hi = tmem_base + F3_H + (uint32_t)(g * 64), a_lo = a_hi + 32u;
          const uint64_t w_hi = ptx::umma_desc_sw128(uW2 + q * F3_WBLK), w_lo = ptx::umma_desc_sw128(uW2 + q * F3_WBLK + 64 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ptx::mma_bf16_ts(d, a_lo + (uint32_t)(k * 8), w_hi + ko, IDESC, (q == 0 && k == 0) ? 0u : 1u);
            ptx::mma_bf16_ts(d, a_hi + (uint32_t)(k * 8), w_lo + ko, IDESC, 1u);
            ptx::mma_bf16_ts(d, a_hi + (uint32_t)(k * 8), w_hi + ko, IDESC, 1u);
          }
          ptx::tc_commit(&h_empty[g]);
          if (q == 3) ptx::tc_commit(&acc2_full[ab]);
        }
      }
    }
  } else {
    // ================= copy warp: resident weights once, then the raw rows of every tile (swizzled 16-byte chunks) =================
    if (my_tiles > 0) {
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(&w_full, 8 * F3_WBLK);
        for (int q = 0; q < 4; ++q) ptx::bulk_g2s(ptx::smem_u32(sW1) + q * F3_WBLK, a.w1 + (size_t)q * F3_WBLK, F3_WBLK, &w_full);
        for (int q = 0; q < 4; ++q) ptx::bulk_g2s(ptx::smem_u32(sW2) + q * F3_WBLK, a.w2 + (size_t)q * F3_WBLK, F3_WBLK, &w_full);
      }
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
        const int s = it % F3_NSLOT;
        ptx::mbar_wait(&x_empty[s], ((uint32_t)(it / F3_NSLOT) & 1u) ^ 1u);
        const uint32_t dst0 = ptx::smem_u32(sX) + s * F3_XSLOT;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
          const int i = k * 32 + lane, r = i >> 4, c = i & 15;
          const int m = m0 + r;
          const int mc = m < a.M ? m : a.M - 1;
          ptx::cp_async16_zfill(dst0 + r * 256 + ((c ^ (r & 7)) << 4), a.x + (long long)mc * 64 + c * 4, m < a.M ? 16u : 0u);
        }
        ptx::cp_async_mbar_arrive(&x_full[s]);
      }
    }
  }
  __syncthreads();
  if (warp == F3_W_MMA1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_ffn_fused(const SebFfn* f, void* stream) {
  SEB_REQUIRE(f && f->x && f->out && f->tokens > 0 && f->tokens < 2147483647LL - 128, SEB_EINVAL, "ffn_fused: bad arguments");
  SEB_REQUIRE(f->ln_gamma && f->ln_beta && f->w1_tc && f->b1 && f->w2_tc && f->b2, SEB_EINVAL, "ffn_fused: null parameter");
  SEB_REQUIRE(aligned16(f->x) && aligned16(f->out) && aligned16(f->w1_tc) && aligned16(f->w2_tc) && aligned16(f->b1) && aligned16(f->b2), SEB_EALIGN, "ffn_fused: unaligned pointer");
  SEB_REQUIRE(aligned16(f->ln_gamma) && aligned16(f->ln_beta), SEB_EALIGN, "ffn_fused: unaligned LayerNorm parameters");
  if (f->post_gamma) SEB_REQUIRE(f->post_beta && f->resid2 && aligned16(f->resid2) && aligned16(f->post_gamma) && aligned16(f->post_beta), SEB_EINVAL,
                                 "ffn_fused: post-norm needs 16-byte aligned gamma / beta and resid2");
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM);
    if (e != cudaSuccess) { set_error("ffn_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  FfnArgs a;
  a.x = f->x; a.out = f->out; a.M = (int)f->tokens; a.ln_g = f->ln_gamma; a.ln_b = f->ln_beta;
  a.w1 = reinterpret_cast<const uint8_t*>(f->w1_tc); a.b1 = f->b1;
  a.w2 = reinterpret_cast<const uint8_t*>(f->w2_tc); a.b2 = f->b2;
  a.alpha = f->alpha; a.pn_g = f->post_gamma; a.pn_b = f->post_beta; a.resid2 = f->resid2;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  const long long ntiles = (f->tokens + BM - 1) / BM;
  const unsigned grid = (unsigned)(ntiles < num_sms ? ntiles : num_sms);      // persistent: one CTA per SM
  ffn_fused_kernel<<<grid, F3_THREADS, F3_SMEM, (cudaStream_t)stream>>>(a);
  SEB_CHECK_LAUNCH("ffn_fused_kernel");
  return 0;
}
