// tok_gemm.cu -- persistent tcgen05 GEMM for the token-wise projections of the conformer (K = 64 or 128):
//   q|k|v projection        LayerNorm -> 64 -> 192, fp16 epilogue            (conformer.py:100-101)
//   pointwise conv + GLU    LayerNorm -> 64 -> 256 -> a * sigmoid(b)         (conformer.py:162-166)
//   attention out-proj      64 -> 64 + bias + residual                       (conformer.py:122-124, 203)
//   pointwise conv 2        128 -> 64 + bias + residual                      (conformer.py:169, 204)
// These are HBM-bound (2.7 - 4.2 GB per launch for ~0.1 TFLOP); round-1 profiles had them at 34 - 50 % of the HBM peak
// because at most one tile's loads were in flight per SM (register-staged loaders) and the A operand made a shared-memory
// round trip.  v3 follows the fused feed-forward kernel:
//   copy warp     : the next tiles' raw fp32 rows into a swizzled staging ring (2 - 3 tiles = 64 - 128 KB in flight per SM, no registers
//                   involved); loads the whole weight image once (resident).  Contiguous 64-float rows (q|k|v, GLU, out-projection) arrive by
//                   tensor-map TMA: ONE thread issues two cp.async.bulk.tensor (left / right 128-byte halves of 128 rows, SWIZZLE_128B, rows past M
//                   zero-filled by the hardware) per tile instead of 2048 cp.async with their address arithmetic (12 % of the kernel's
//                   instructions); the other shapes (K = 128, fp16 rows, two sources) keep the cp.async path
//   4 row warps   : thread = row.  staged row -> [LayerNorm] -> bf16 hi|lo -> tcgen05.st into XA[s] (tensor memory)
//   MMA issuer    : ACC[ab] = XA[s] . W^T, A operand from TMEM (`[a_tmem]` form), 3-product split
//   16 epilogue warps : tcgen05.ld -> warp-private smem transpose -> the engine's epilogue functors, coalesced stores
#include "gemm_engine.cuh"
#include <cuda.h>          // CUtensorMap (types only: the encoder comes through cudaGetDriverEntryPoint, libcuda is not linked)
#include <stdlib.h>
#include <string.h>

namespace seb {

constexpr int TG_ROW_WARPS = 4, TG_EPI_WARPS = 16;
constexpr int TG_W_EPI0 = TG_ROW_WARPS;                                       // multiple of 4: warp % 4 = TMEM lane quarter
constexpr int TG_W_MMA = TG_W_EPI0 + TG_EPI_WARPS;                            // 20
constexpr int TG_W_COPY = TG_W_MMA + 1;                                       // 21
constexpr int TG_THREADS = (TG_W_COPY + 1) * 32;                              // 704
// AH: the A rows are stored as __half (SEB_LOAD_ROWS_F16): half the bytes per staged tile
template <int KCH, bool AH = false> constexpr int tg_slots() { return (KCH == 1 || AH) ? 3 : 2; }
template <int KCH, bool AH = false> constexpr int tg_xslot() { return BM * 64 * KCH * (AH ? 2 : 4); }     // raw rows of one tile: 32 / 64 KB (fp16: 16 / 32 KB)
// epilogue staging per warp: the packed epilogues (GLU / fp16 q|k|v / gate) transpose 32 x 64 B, the others 32 x 128 B
template <int EK> constexpr bool tg_packed() { return EK == SEB_EPI_GLU || EK == SEB_EPI_GLU_F16 || EK == SEB_EPI_QKV_F16 || EK == SEB_EPI_GATE; }
template <int EK> constexpr int tg_stg() { return tg_packed<EK>() ? 2048 : 4096; }
template <int NT, int KCH, int EK, bool AH = false> constexpr int tg_smem_bytes() {
  return 1024 + tg_slots<KCH, AH>() * tg_xslot<KCH, AH>() + KCH * 2 * NT * 128 + TG_EPI_WARPS * tg_stg<EK>() + 2 * 64 * 4 + NT * 4;
}

#ifdef TG_TRACE
__device__ long long tg_trace[4][24][4];     // [role: row warp 0, MMA issuer, epilogue warp 0, copy warp][tile][event] clock64 stamps of CTA TG_TRACE (debug builds only)
#define TG_STAMP(role, it, ev) do { if (blockIdx.x == TG_TRACE && (it) < 24 && lane == 0) tg_trace[role][it][ev] = clock64(); } while (0)
#else
#define TG_STAMP(role, it, ev) do { } while (0)
#endif

namespace ptx {
__device__ __forceinline__ void tg_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tg_tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
}  // namespace ptx

// rows of 256 bytes: 64 floats (K = 64) or 128 halfs (K = 128, fp16 rows)
template <int KCH, int LK> constexpr bool tg_tma() { return (KCH == 1 && (LK == SEB_LOAD_ROWS || LK == SEB_LOAD_ROWS_LN)) || (KCH == 2 && LK == SEB_LOAD_ROWS_F16); }

template <int NT, int KCH, int LK, int EK, bool TMA = false>
__global__ void __launch_bounds__(TG_THREADS, 1) tok_gemm_kernel(const GemmArgs g, const uint8_t* __restrict__ w_tc, const __grid_constant__ CUtensorMap tmx) {
  static_assert(!TMA || tg_tma<KCH, LK>(), "the TMA staging path covers contiguous 64-float rows");
  static_assert(NT % 64 == 0 && NT <= 256 && (KCH == 1 || KCH == 2), "unsupported token GEMM shape");
  static_assert(LK == SEB_LOAD_ROWS || LK == SEB_LOAD_ROWS_F16 || (LK == SEB_LOAD_ROWS_LN && KCH == 1) || (LK == SEB_LOAD_ROWS2 && KCH == 2),
                "loader: plain rows (fp32 or fp16), LayerNorm over 64 features, or two 64-wide sources side by side");
  constexpr bool AH = LK == SEB_LOAD_ROWS_F16;
  constexpr int STG = tg_stg<EK>();
  constexpr int K = 64 * KCH, NSLOT = tg_slots<KCH, AH>(), XSLOT = tg_xslot<KCH, AH>(), PITCH = K * (AH ? 2 : 4), NCH = PITCH / 16;   // 16-byte chunks per row
  constexpr bool ACC2 = (2 * K + 2 * NT <= 512);                 // double-buffered accumulator when tensor memory has room
  constexpr uint32_t T_ACC = 2 * K;                              // TMEM: XA[2] (K columns each: hi | lo) | ACC[1 or 2] (NT columns each)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t x_full[NSLOT], x_empty[NSLOT], xa_full[2], xa_empty[2], w_full, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                   // [kc][hi | lo][NT rows x 128 B], resident
  uint8_t* sX = sW + KCH * 2 * NT * 128;                // ring of raw fp32 tiles, 16-byte chunk c of row r at r * PITCH + ((c ^ (r & 7)) << 4)
  uint8_t* sStg = sX + NSLOT * XSLOT;                   // 4 KB per epilogue warp
  float* sG = reinterpret_cast<float*>(sStg + TG_EPI_WARPS * STG);
  float* sBt = sG + 64;
  float* sBias = sBt + 64;                               // [NT] output bias (zeros when the GEMM has none)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (g.M + BM - 1) / BM;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) { ptx::mbar_init(&x_full[i], TMA ? 1 : 32); ptx::mbar_init(&x_empty[i], TG_ROW_WARPS * 32); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&xa_full[i], TG_ROW_WARPS * 32); ptx::mbar_init(&xa_empty[i], 1);
      ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], TG_EPI_WARPS * 32);
    }
    ptx::mbar_init(&w_full, 1);
    ptx::fence_barrier_init();
  }
  if (LK == SEB_LOAD_ROWS_LN && tid < 128) { if (tid < 64) sG[tid] = g.ln_g[tid]; else sBt[tid - 64] = g.ln_b[tid - 64]; }
  for (int i = tid; i < NT; i += TG_THREADS) sBias[i] = (g.bias && i < g.N) ? g.bias[i] : 0.f;
  if (warp == TG_W_MMA) ptx::tmem_alloc(&tmem_base_s, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < TG_ROW_WARPS) {
    // ================= row warps: thread = row; staged row -> [LayerNorm] -> bf16 hi|lo -> XA[s] (TMEM) =================
    const int row = tid, sw = row & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int it = 0; it < my_tiles; ++it) {
      const int slot = it % NSLOT, s2 = it & 1;
      // TMA layout: two [128 rows x 128 B] halves per slot, 16-byte chunk XOR (row & 7) inside each 128-byte row (SWIZZLE_128B)
      const uint8_t* xr = sX + slot * XSLOT + row * (TMA ? 128 : PITCH);
      auto chunk = [&](int c) -> const uint8_t* { return TMA ? xr + ((c >> 3) << 14) + (((c & 7) ^ sw) << 4) : xr + ((c ^ sw) << 4); };
      if (warp == 0) TG_STAMP(0, it, 0);
      ptx::mbar_wait(&x_full[slot], (uint32_t)(it / NSLOT) & 1u);
      if (warp == 0) TG_STAMP(0, it, 1);
      float mean = 0.f, rstd = 1.f;
      if (LK == SEB_LOAD_ROWS_LN) {     // shifted one-pass statistics, four partial sums (short dependency chains)
        float2 ps[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, pq[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        const float x0 = reinterpret_cast<const float4*>(chunk(0))->x;
        const float2 nx0 = make_float2(-x0, -x0);
#pragma unroll
        for (int c = 0; c < NCH; ++c) stats_acc4(*reinterpret_cast<const float4*>(chunk(c)), nx0, ps[c & 1], pq[c & 1]);
        const float md = ((ps[0].x + ps[0].y) + (ps[1].x + ps[1].y)) * (1.0f / 64.0f);
        const float var = fmaxf(((pq[0].x + pq[0].y) + (pq[1].x + pq[1].y)) * (1.0f / 64.0f) - md * md, 0.f);
        rstd = 1.0f / sqrtf(var + 1e-5f);
        mean = x0 + md;
      }
      ptx::mbar_wait(&xa_empty[s2], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      ptx::tc_fence_after();
      if (warp == 0) TG_STAMP(0, it, 2);
      const uint32_t xa = lane_base + (uint32_t)(s2 * K);
#pragma unroll
      for (int c16 = 0; c16 < K / 16; ++c16) {          // 16 k-values -> 8 hi + 8 lo packed columns
        uint32_t hi[8], lo[8];
        if (AH) {                                       // 16 halfs = two 16-byte chunks
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = c16 * 2 + j;
            const uint4 h = *reinterpret_cast<const uint4*>(chunk(c));
            const uint32_t w4[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[i]));
              split_bf16x2(f.x, f.y, hi[4 * j + i], lo[4 * j + i]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = c16 * 4 + j;
            const float4 v = *reinterpret_cast<const float4*>(chunk(c));
            float2 y01 = make_float2(v.x, v.y), y23 = make_float2(v.z, v.w);
            if (LK == SEB_LOAD_ROWS_LN)
              ln_apply4(v, mean, rstd, *reinterpret_cast<const float4*>(sG + c * 4), *reinterpret_cast<const float4*>(sBt + c * 4), y01, y23);
            split_bf16x2(y01.x, y01.y, hi[2 * j], lo[2 * j]);
            split_bf16x2(y23.x, y23.y, hi[2 * j + 1], lo[2 * j + 1]);
          }
        }
        ptx::tg_tmem_st8(xa + (uint32_t)(c16 * 8), hi);
        ptx::tg_tmem_st8(xa + (uint32_t)(K / 2 + c16 * 8), lo);
      }
      ptx::mbar_arrive(&x_empty[slot]);          // the staged row is consumed: the copy warp may refill the slot
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ptx::tc_fence_before();
      ptx::mbar_arrive(&xa_full[s2]);
      if (warp == 0) TG_STAMP(0, it, 3);
    }
  } else if (warp < TG_W_MMA) {
    // ================= epilogue warps: TMEM -> warp-private smem transpose -> coalesced functor =================
    const int ew = warp - TG_W_EPI0;
    const int wq = warp & 3, cgi = ew >> 2;             // TMEM lane quarter (hardware: warp % 4), column group
    constexpr int CPW = NT / (TG_EPI_WARPS / 4);        // columns per warp (64 / 48 / 16)
    float4* stg = reinterpret_cast<float4*>(sStg + ew * STG);       // [32 rows][8 x float4] (packed epilogues: [32 rows][4 x 16 B])
    // the bias of this lane's columns is loop invariant: keep it in registers and hand the functors a bias-free descriptor
    // (inside Epi<>::apply the load sat in front of every use: ~20 % of the epilogue's stall samples)
    GemmArgs gnb = g;
    gnb.bias = nullptr;
    float4 bias4[(CPW + 31) / 32];
#pragma unroll
    for (int c0 = 0; c0 < CPW; c0 += 32) {
      const int nb = cgi * CPW + c0 + (lane & 7) * 4;
      bias4[c0 / 32] = (g.bias && (lane & 7) * 4 < CPW - c0 && nb + 4 <= g.N) ? ldg4(g.bias + nb) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
      const int ab = ACC2 ? (it & 1) : 0;
      const uint32_t use = ACC2 ? (uint32_t)(it >> 1) : (uint32_t)it;
      const int ch = lane & 7;
      // residual rows of this warp's outputs: issued BEFORE the accumulator wait (out may alias resid, so the compiler
      // cannot hoist them over the stores of the copy-out loop itself; serialised they cost eight DRAM round trips per tile)
      float4 res[8];
      if (EK == SEB_EPI_RESID || EK == SEB_EPI_RESID_SCALE) {
        static_assert((EK != SEB_EPI_RESID && EK != SEB_EPI_RESID_SCALE) || CPW <= 32, "residual prefetch covers one 32-column pass");
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int m = m0 + wq * 32 + i8 * 4 + (lane >> 3), n = cgi * CPW + ch * 4;
          res[i8] = (ch * 4 < CPW && m < g.M) ? *reinterpret_cast<const float4*>(g.resid + (long long)m * g.ldr + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (ew == 0) TG_STAMP(2, it, 0);
      ptx::mbar_wait(&acc_full[ab], use & 1u);
      ptx::tc_fence_after();
      if (ew == 0) TG_STAMP(2, it, 1);
      if (ew == 0 && it > 0) TG_STAMP(2, it - 1, 3);
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + T_ACC + (uint32_t)(ab * NT + cgi * CPW);
      if (tg_packed<EK>()) {
        // element-wise part in the accumulator's own thread = row layout (bias, GLU / fp16 scaling), THEN the transpose: the
        // staged tile and the copy-out carry the packed outputs only (half the words of the fp32 accumulator columns)
        uint4* stq = reinterpret_cast<uint4*>(stg);                   // [32 rows][4 x 16 B], chunk XOR ((row >> 1) & 3)
#pragma unroll
        for (int c0 = 0; c0 < CPW; c0 += 32) {
          constexpr int dummy = 0; (void)dummy;
          const int ncols = (CPW - c0 < 32) ? CPW - c0 : 32;
          const int n0 = cgi * CPW + c0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
              float t8[8];
              ptx::tmem_ld8(taddr + c0 + j, t8);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[j + i] = t8[i];
            }
          }
          if (c0 + 32 >= CPW) { ptx::tc_fence_before(); ptx::mbar_arrive(&acc_empty[ab]); if (ew == 0) TG_STAMP(2, it, 2); }     // accumulator fully read
          const float* rb = nullptr;               // gate: this row's group bias (diffusion-step projection), L1-resident
          if (EK == SEB_EPI_GATE && g.resid) {
            const int mr = m0 + wq * 32 + lane;
            rb = g.resid + (long long)((mr < g.M ? mr : g.M - 1) / (int)g.ldr) * g.N + n0;
          }
          if (EK == SEB_EPI_GLU_F16) {             // 32 accumulator columns -> 16 halfs = two 16-byte chunks
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              if (qq * 16 < ncols) {
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int c = 16 * qq + 4 * i;
                  const float4 b = *reinterpret_cast<const float4*>(sBias + n0 + c);
                  const __half2 h = __floats2half2_rn((v[c + 0] + b.x) * sigmoidf_acc(v[c + 1] + b.y), (v[c + 2] + b.z) * sigmoidf_acc(v[c + 3] + b.w));
                  pk[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                stq[lane * 4 + (qq ^ ((lane >> 1) & 3))] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
            __syncwarp();
            const int cc = lane & 3;
            if (cc * 16 < ncols) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int R = 8 * i + (lane >> 2);
                const int m = m0 + wq * 32 + R;
                const uint4 o = stq[R * 4 + (cc ^ ((R >> 1) & 3))];
                if (m < g.M) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(g.out) + (long long)m * g.ldo + (n0 >> 1) + cc * 8) = o;
              }
            }
            __syncwarp();
            continue;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {            // output chunk q <- accumulator columns 8q .. 8q + 7
            if (q * 8 < ncols) {
              uint4 o;
              if (EK == SEB_EPI_GLU) {             // packed columns: (value_j, gate_j) adjacent (conformer.py:36-37)
                const float4 b0 = *reinterpret_cast<const float4*>(sBias + n0 + 8 * q), b1 = *reinterpret_cast<const float4*>(sBias + n0 + 8 * q + 4);
                o.x = __float_as_uint((v[8 * q + 0] + b0.x) * sigmoidf_acc(v[8 * q + 1] + b0.y));
                o.y = __float_as_uint((v[8 * q + 2] + b0.z) * sigmoidf_acc(v[8 * q + 3] + b0.w));
                o.z = __float_as_uint((v[8 * q + 4] + b1.x) * sigmoidf_acc(v[8 * q + 5] + b1.y));
                o.w = __float_as_uint((v[8 * q + 6] + b1.z) * sigmoidf_acc(v[8 * q + 7] + b1.w));
              } else if (EK == SEB_EPI_GATE) {     // packed columns: (gate_j, filter_j) adjacent (tsc_diffusion.py:36-37)
                float4 b0 = *reinterpret_cast<const float4*>(sBias + n0 + 8 * q), b1 = *reinterpret_cast<const float4*>(sBias + n0 + 8 * q + 4);
                if (rb) {
                  const float4 r0 = ldg4(rb + 8 * q), r1 = ldg4(rb + 8 * q + 4);
                  b0.x += r0.x; b0.y += r0.y; b0.z += r0.z; b0.w += r0.w; b1.x += r1.x; b1.y += r1.y; b1.z += r1.z; b1.w += r1.w;
                }
                o.x = __float_as_uint(sigmoidf_acc(v[8 * q + 0] + b0.x) * tanhf_acc(v[8 * q + 1] + b0.y));
                o.y = __float_as_uint(sigmoidf_acc(v[8 * q + 2] + b0.z) * tanhf_acc(v[8 * q + 3] + b0.w));
                o.z = __float_as_uint(sigmoidf_acc(v[8 * q + 4] + b1.x) * tanhf_acc(v[8 * q + 5] + b1.y));
                o.w = __float_as_uint(sigmoidf_acc(v[8 * q + 6] + b1.z) * tanhf_acc(v[8 * q + 7] + b1.w));
              } else {                             // fp16 q | k | v, q pre-scaled by dim_head^-0.5 * log2(e)
                const float sc = (n0 + 8 * q < 64) ? 0.25f * 1.4426950408889634f : 1.0f;
                __half2 h0 = __floats2half2_rn(v[8 * q + 0] * sc, v[8 * q + 1] * sc), h1 = __floats2half2_rn(v[8 * q + 2] * sc, v[8 * q + 3] * sc);
                __half2 h2 = __floats2half2_rn(v[8 * q + 4] * sc, v[8 * q + 5] * sc), h3 = __floats2half2_rn(v[8 * q + 6] * sc, v[8 * q + 7] * sc);
                o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              }
              stq[lane * 4 + (q ^ ((lane >> 1) & 3))] = o;
            }
          }
          __syncwarp();
          const int cc = lane & 3;
          if (cc * 8 < ncols) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int R = 8 * i + (lane >> 2);
              const int m = m0 + wq * 32 + R;
              const uint4 o = stq[R * 4 + (cc ^ ((R >> 1) & 3))];
              if (m < g.M) {
                if (EK == SEB_EPI_GLU || EK == SEB_EPI_GATE) *reinterpret_cast<uint4*>(g.out + (long long)m * g.ldo + (n0 >> 1) + cc * 4) = o;
                else *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(g.out) + (long long)m * g.ldo + n0 + cc * 8) = o;
              }
            }
          }
          __syncwarp();
        }
        continue;
      }
#pragma unroll
      for (int c0 = 0; c0 < CPW; c0 += 32) {
        const int ncols = (CPW - c0 < 32) ? CPW - c0 : 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncols) {
            float v[8];
            ptx::tmem_ld8(taddr + c0 + j, v);
            stg[lane * 8 + (((j >> 2) + 0) ^ (lane & 7))] = make_float4(v[0], v[1], v[2], v[3]);
            stg[lane * 8 + (((j >> 2) + 1) ^ (lane & 7))] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
        if (c0 + 32 >= CPW) { ptx::tc_fence_before(); ptx::mbar_arrive(&acc_empty[ab]); }     // accumulator fully read
        __syncwarp();
        float4 vals[8];
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {          // all staging reads first: nothing between them and their use can alias
          const int R = i8 * 4 + (lane >> 3);
          vals[i8] = stg[R * 8 + (ch ^ (R & 7))];
        }
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int R = i8 * 4 + (lane >> 3);
          if (ch * 4 < ncols) {
            const int m = m0 + wq * 32 + R, n = cgi * CPW + c0 + ch * 4;
            if (EK == SEB_EPI_RESID) {
              if (m < g.M && n < g.N) {
                float4 v = vals[i8];
                const float4 bb = bias4[c0 / 32];
                v.x = fmaf(g.alpha, v.x + bb.x, res[i8].x); v.y = fmaf(g.alpha, v.y + bb.y, res[i8].y);
                v.z = fmaf(g.alpha, v.z + bb.z, res[i8].z); v.w = fmaf(g.alpha, v.w + bb.w, res[i8].w);
                st4(g.out + (long long)m * g.ldo + n, v);
              }
            } else if (EK == SEB_EPI_RESID_SCALE) {
              if (m < g.M && n < g.N) {
                float4 v = vals[i8];
                const float4 bb = bias4[c0 / 32];
                v.x = g.alpha * (v.x + bb.x + res[i8].x); v.y = g.alpha * (v.y + bb.y + res[i8].y);
                v.z = g.alpha * (v.z + bb.z + res[i8].z); v.w = g.alpha * (v.w + bb.w + res[i8].w);
                st4(g.out + (long long)m * g.ldo + n, v);
              }
            } else {
              float4 v = vals[i8];
              const float4 bb = bias4[c0 / 32];
              v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              Epi<EK>::apply(gnb, m, n, v);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == TG_W_MMA) {
    // ================= MMA issuer: ACC[ab] = XA[s] . W^T (A from tensor memory) =================
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t uW = ptx::smem_u32(sW);
      ptx::mbar_wait(&w_full, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int s2 = it & 1;
        const int ab = ACC2 ? (it & 1) : 0;
        const uint32_t use = ACC2 ? (uint32_t)(it >> 1) : (uint32_t)it;
        TG_STAMP(1, it, 0);
        ptx::mbar_wait(&xa_full[s2], (uint32_t)(it >> 1) & 1u);
        TG_STAMP(1, it, 1);
        ptx::mbar_wait(&acc_empty[ab], (use & 1u) ^ 1u);
        ptx::tc_fence_after();
        TG_STAMP(1, it, 2);
        const uint32_t d_tmem = tmem_base + T_ACC + (uint32_t)(ab * NT);
        const uint32_t a_hi0 = tmem_base + (uint32_t)(s2 * K), a_lo0 = a_hi0 + (uint32_t)(K / 2);
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
          const uint64_t w_hi = ptx::umma_desc_sw128(uW + kc * 2 * NT * 128), w_lo = ptx::umma_desc_sw128(uW + kc * 2 * NT * 128 + NT * 128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            const uint32_t ac = (uint32_t)(kc * 32 + k * 8);
            ptx::tg_mma_ts(d_tmem, a_lo0 + ac, w_hi + ko, IDESC, (kc | k) ? 1u : 0u);
            ptx::tg_mma_ts(d_tmem, a_hi0 + ac, w_lo + ko, IDESC, 1u);
            ptx::tg_mma_ts(d_tmem, a_hi0 + ac, w_hi + ko, IDESC, 1u);
          }
        }
        ptx::tc_commit(&xa_empty[s2]);
        ptx::tc_commit(&acc_full[ab]);
        TG_STAMP(1, it, 3);
      }
    }
  } else {
    // ================= copy warp: weights once, then the raw rows of every tile (swizzled 16-byte chunks) =================
    if (my_tiles > 0) {
      if (lane == 0) {
        constexpr uint32_t WB = KCH * 2 * NT * 128;
        ptx::mbar_arrive_expect_tx(&w_full, WB);
        constexpr uint32_t PIECE = 16384;                  // bulk copies of 16 KB
        for (uint32_t o = 0; o < WB; o += PIECE) ptx::bulk_g2s(ptx::smem_u32(sW) + o, w_tc + o, (WB - o < PIECE) ? WB - o : PIECE, &w_full);
      }
      const float* src0 = g.a[0];
      const float* src1 = (LK == SEB_LOAD_ROWS2) ? g.a[1] : nullptr;
      if (TMA) {
        if (lane == 0) {
          asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
          for (int it = 0; it < my_tiles; ++it) {
            const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
            const int slot = it % NSLOT;
            TG_STAMP(3, it, 0);
            ptx::mbar_wait(&x_empty[slot], ((uint32_t)(it / NSLOT) & 1u) ^ 1u);
            TG_STAMP(3, it, 1);
            const uint32_t dst0 = ptx::smem_u32(sX) + slot * XSLOT, bar = ptx::smem_u32(&x_full[slot]);
            ptx::mbar_arrive_expect_tx(&x_full[slot], (uint32_t)XSLOT);
#pragma unroll
            for (int h = 0; h < 2; ++h)
              asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                           ::"r"(dst0 + (uint32_t)(h << 14)), "l"(reinterpret_cast<uint64_t>(&tmx)), "r"(0), "r"(h), "r"(m0), "r"(bar) : "memory");
            TG_STAMP(3, it, 2);
          }
        }
      } else
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
        const int slot = it % NSLOT;
        TG_STAMP(3, it, 0);
        ptx::mbar_wait(&x_empty[slot], ((uint32_t)(it / NSLOT) & 1u) ^ 1u);
        TG_STAMP(3, it, 1);
        const uint32_t dst0 = ptx::smem_u32(sX) + slot * XSLOT;
#pragma unroll 8
        for (int kk = 0; kk < BM * NCH / 32; ++kk) {
          const int i = kk * 32 + lane, r = i / NCH, c = i % NCH;
          const int m = m0 + r;
          const int mc = m < g.M ? m : g.M - 1;
          const void* src = (LK == SEB_LOAD_ROWS2) ? (const void*)((c < 16 ? src0 : src1) + (long long)mc * g.lda + (c & 15) * 4)
                            : AH ? (const void*)(reinterpret_cast<const __half*>(src0) + (long long)mc * g.lda + c * 8)
                                 : (const void*)(src0 + (long long)mc * g.lda + c * 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                       ::"r"(dst0 + r * PITCH + ((c ^ (r & 7)) << 4)), "l"(src), "r"(m < g.M ? 16u : 0u) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(&x_full[slot])) : "memory");
        TG_STAMP(3, it, 2);
      }
    }
  }
  __syncthreads();
  if (warp == TG_W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

#ifdef TG_TRACE
extern "C" int seb200_tg_trace(long long* host) { return (int)cudaMemcpyFromSymbol(host, tg_trace, sizeof(tg_trace)); }
#endif

typedef CUresult (*tg_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tg_encode_fn tg_encoder() {
  static tg_encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<tg_encode_fn>(p);
  }();
  return fn;
}
// x [M, 64] fp32, contiguous rows, as (32 floats, 2 halves of a row, M rows); box = one 128-byte half of 128 rows, SWIZZLE_128B
static bool tg_make_map(CUtensorMap* tm, const float* x, int M) {
  tg_encode_fn enc = tg_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {32, 2, (cuuint64_t)M};
  const cuuint64_t strides[2] = {128, 256};
  const cuuint32_t box[3] = {32, 1, (cuuint32_t)BM}, estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NT, int KCH, int LK, int EK>
static int launch_tok(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  static int num_sms = 0;
  static const bool no_tma = getenv("SEB200_TOK_NO_TMA") && atoi(getenv("SEB200_TOK_NO_TMA")) != 0;      // A/B switch
  constexpr int SMEM = tg_smem_bytes<NT, KCH, EK, LK == SEB_LOAD_ROWS_F16>();
  static_assert(SMEM + 256 <= 232448, "token GEMM: shared memory over the 227 KB per-CTA limit");
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(tok_gemm_kernel<NT, KCH, LK, EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (tg_tma<KCH, LK>() && e == cudaSuccess) e = cudaFuncSetAttribute(tok_gemm_kernel<NT, KCH, LK, EK, tg_tma<KCH, LK>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("tok gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    attr_done.set();
  }
  const long long ntiles = ((long long)s->M + BM - 1) / BM;
  dim3 grid((unsigned)(ntiles < num_sms ? ntiles : num_sms));
  CUtensorMap tm;
  if (tg_tma<KCH, LK>() && !no_tma && g.lda == (LK == SEB_LOAD_ROWS_F16 ? 128 : 64) && aligned16(g.a[0]) && tg_make_map(&tm, g.a[0], g.M)) {
    tok_gemm_kernel<NT, KCH, LK, EK, tg_tma<KCH, LK>()><<<grid, TG_THREADS, SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc), tm);
    SEB_CHECK_LAUNCH("tok_gemm_kernel<tma>");
    return 0;
  }
  memset(&tm, 0, sizeof(tm));
  tok_gemm_kernel<NT, KCH, LK, EK><<<grid, TG_THREADS, SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc), tm);
  SEB_CHECK_LAUNCH("tok_gemm_kernel");
  return 0;
}

// returns -100 when the shape is not one of the persistent token GEMMs (caller falls back to the tile-per-CTA engine)
int launch_tok_gemm(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("SEB200_TOK_PERSIST"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled || s->tc_planes != 2 || s->tc_ntiles != 1) return -100;
  const int nt = s->tc_ntile;
  if (s->loader == SEB_LOAD_ROWS_LN && s->epilogue == SEB_EPI_QKV_F16 && nt == 192 && s->K == 64)
    return launch_tok<192, 1, SEB_LOAD_ROWS_LN, SEB_EPI_QKV_F16>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS_LN && s->epilogue == SEB_EPI_GLU && nt == 256 && s->K == 64)
    return launch_tok<256, 1, SEB_LOAD_ROWS_LN, SEB_EPI_GLU>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS_LN && s->epilogue == SEB_EPI_GLU_F16 && nt == 256 && s->K == 64)
    return launch_tok<256, 1, SEB_LOAD_ROWS_LN, SEB_EPI_GLU_F16>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS_F16 && s->epilogue == SEB_EPI_RESID && nt == 64 && s->K == 128 && s->lda == 128)
    return launch_tok<64, 2, SEB_LOAD_ROWS_F16, SEB_EPI_RESID>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS && s->epilogue == SEB_EPI_RESID && nt == 64 && s->K == 64 && s->lda == 64)
    return launch_tok<64, 1, SEB_LOAD_ROWS, SEB_EPI_RESID>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS && s->epilogue == SEB_EPI_RESID && nt == 64 && s->K == 128 && s->lda == 128)
    return launch_tok<64, 2, SEB_LOAD_ROWS, SEB_EPI_RESID>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS2 && s->epilogue == SEB_EPI_GATE && nt == 128 && s->K == 128 && s->N == 128)
    return launch_tok<128, 2, SEB_LOAD_ROWS2, SEB_EPI_GATE>(s, g, st);
  if (s->loader == SEB_LOAD_ROWS && s->epilogue == SEB_EPI_RESID_SCALE && nt == 64 && s->K == 64 && s->lda == 64)
    return launch_tok<64, 1, SEB_LOAD_ROWS, SEB_EPI_RESID_SCALE>(s, g, st);
  return -100;
}

}  // namespace seb
