// ffn_fused.cu -- one kernel for the conformer feed-forward half-step (conformer.py:128-145 wrapped by PreNorm and
// Scale(0.5), conformer.py:53-71, 201-202, 207, 210), optionally followed by post_norm and the TSCB outer residual
// (conformer.py:211, generator.py:70,72):
//
//     y   = x + alpha * ( W2 . swish( W1 . LayerNorm(x) + b1 ) + b2 )
//     out = post ? LayerNorm_post(y) + resid2 : y
//
// The 128 x 256 hidden activation never leaves the SM.  Per 128-token CTA tile:
//   producers (8 warps)  : x -> LayerNorm -> bf16 hi/lo -> swizzled smem A operand
//   for q in 0..3        : MMA1(q): acc1[q&1] (TMEM, 64 cols) = A . W1[64q:64q+64]^T          (tcgen05, 3-product split)
//                          epilogue warps: tcgen05.ld acc1 -> +b1 -> swish -> bf16 hi/lo -> smem H (K-chunk q of GEMM 2)
//                          MMA2(q): acc2 (TMEM, 64 cols) += H . W2[:, 64q:64q+64]^T
//   final                : tcgen05.ld acc2 -> smem transpose -> coalesced: *alpha + b2 + x (-> LayerNorm + resid2) -> store
// acc1 is double buffered, so MMA1(q+1) overlaps the Swish epilogue of chunk q.
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

// ---- persistent, warp-specialised version ----------------------------------------------------------------------
// One CTA per SM loops over 128-token tiles.  W1 (hi|lo, 64 KB) is loaded ONCE and stays resident in shared memory, W2
// blocks stream through a 2-slot ring.  The raw fp32 rows of the NEXT tile (128 x 256 B, contiguous in HBM) are staged
// into shared memory by one cp.async.bulk issued by a dedicated warp, so the 4 loader warps (LayerNorm + bf16 hi/lo
// split into the double-buffered A operand) never wait on a global load: round 1 profiling showed the loaders, with
// eight serialised ~740-cycle LDG waits per tile, pacing the whole kernel.  The hidden chunk H is single buffered
// (its writer only needs it after its MUFU phase, by which time MMA2 of the previous chunk has long retired);
// 16 epilogue warps do the Swish mid-epilogues and the final store; 1 thread issues every tcgen05.mma.
// acc1 and acc2 are double buffered in TMEM, so MMA1 of tile i+1 overlaps the final epilogue of tile i.
constexpr int FF_LOAD_WARPS = 4, FF_EPI_WARPS = 16;
constexpr int FF_EPI_THREADS = FF_EPI_WARPS * 32;
constexpr int FF_CG = FF_EPI_WARPS / 4;            // column groups: each epilogue thread owns one row x (64 / FF_CG) columns
constexpr int FF_CPT = 64 / FF_CG;                 // columns per thread per quarter (16)
constexpr int FF_THREADS = (FF_LOAD_WARPS + FF_EPI_WARPS + 3) * 32;     // + MMA warp + weight warp + x-staging warp = 736
constexpr int FF_PLANE = BM * 128;                 // 16 KB: one bf16 plane of a 128 x 64 operand tile
constexpr int FF_WBLK = 2 * 64 * 128;              // 16 KB: hi|lo image of a 64-row x 64-k weight block
constexpr int FF_XBYTES = BM * 64 * 4;             // 32 KB: raw fp32 rows of one tile
constexpr int FF_SMEM = 1024 + 2 * (2 * FF_PLANE) /*A x2*/ + 4 * FF_WBLK /*W1 resident*/ + 2 * FF_WBLK /*W2 ring*/ + 2 * FF_PLANE /*H*/ + FF_XBYTES /*x staging*/;

__global__ void __launch_bounds__(FF_THREADS, 1) ffn_fused_kernel(const FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[2], a_empty[2], w_full, w2_full[2], w2_empty[2], acc1_full[2], acc1_empty[2], h_full, h_empty, x_full, x_empty, acc2_full[2], acc2_empty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS/LDS, not generic ST/LD)
  uint8_t* sA = smem;                           // [2][hi | lo]            64 KB
  uint8_t* sW1 = sA + 4 * FF_PLANE;             // 4 blocks x (hi | lo)    64 KB  resident
  uint8_t* sW2 = sW1 + 4 * FF_WBLK;             // 2-slot ring of (hi | lo) blocks   32 KB  (streamed: 64 KB per tile from L2)
  uint8_t* sH = sW2 + 2 * FF_WBLK;              // [hi | lo]               32 KB  (doubles as the fp32 staging tile of the final store)
  uint8_t* sX = sH + 2 * FF_PLANE;              // raw fp32 rows of the tile the loaders work on   32 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (a.M + BM - 1) / BM;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc1_full[i], 1); ptx::mbar_init(&acc1_empty[i], FF_EPI_WARPS * 32);
      ptx::mbar_init(&acc2_full[i], 1); ptx::mbar_init(&acc2_empty[i], FF_EPI_WARPS * 32);
    }
    ptx::mbar_init(&h_full, FF_EPI_WARPS * 32); ptx::mbar_init(&h_empty, 1);
    ptx::mbar_init(&x_full, 1); ptx::mbar_init(&x_empty, FF_LOAD_WARPS * 32);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], FF_LOAD_WARPS * 32); ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&w2_full[i], 1); ptx::mbar_init(&w2_empty[i], 1);
    }
    ptx::mbar_init(&w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == FF_LOAD_WARPS + FF_EPI_WARPS) ptx::tmem_alloc(&tmem_base_s, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;     // acc1[b] at column 64 b, acc2[b] at column 128 + 64 b

  if (warp < FF_LOAD_WARPS) {
    // ================= loaders: staged x rows -> LayerNorm -> bf16 hi/lo -> swizzled A[s] =================
    const int sub = tid & 7, rloc = tid >> 3;          // 16 rows per pass, 8 passes; 8 adjacent lanes own one row
    const float4 g0 = ldg4(a.ln_g + sub * 8), g1 = ldg4(a.ln_g + sub * 8 + 4);
    const float4 b0 = ldg4(a.ln_b + sub * 8), b1 = ldg4(a.ln_b + sub * 8 + 4);
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
      // pull the residual rows of the NEXT tile (read by the final epilogue) into L2 while this tile is processed
      if (it + 1 < my_tiles && a.resid2 && a.resid2 != a.x) {
        const long long nrow = (long long)m0 + (long long)gridDim.x * BM + tid;
        if (nrow < (long long)a.M) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.resid2 + nrow * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.resid2 + nrow * 64 + 32));
        }
      }
      const int s = it & 1;
      ptx::mbar_wait(&x_full, (uint32_t)it & 1u);
      ptx::mbar_wait(&a_empty[s], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      uint8_t* dA = sA + s * 2 * FF_PLANE;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int r = p * 16 + rloc;
        float v[8];
        {
          const float4 x0 = *reinterpret_cast<const float4*>(sX + r * 256 + sub * 32);
          const float4 x1 = *reinterpret_cast<const float4*>(sX + r * 256 + sub * 32 + 16);
          v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
        }
        if (m0 + r >= a.M) {        // rows past the end of the tensor were not staged: keep the operand finite
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float sm = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        sm += __shfl_xor_sync(0xffffffffu, sm, 1);
        sm += __shfl_xor_sync(0xffffffffu, sm, 2);
        sm += __shfl_xor_sync(0xffffffffu, sm, 4);
        const float mean = sm * (1.0f / 64.0f);
        float qv = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] -= mean; qv = fmaf(v[i], v[i], qv); }
        qv += __shfl_xor_sync(0xffffffffu, qv, 1);
        qv += __shfl_xor_sync(0xffffffffu, qv, 2);
        qv += __shfl_xor_sync(0xffffffffu, qv, 4);
        const float rstd = 1.0f / sqrtf(qv * (1.0f / 64.0f) + 1e-5f);
        v[0] = v[0] * rstd * g0.x + b0.x; v[1] = v[1] * rstd * g0.y + b0.y;
        v[2] = v[2] * rstd * g0.z + b0.z; v[3] = v[3] * rstd * g0.w + b0.w;
        v[4] = v[4] * rstd * g1.x + b1.x; v[5] = v[5] * rstd * g1.y + b1.y;
        v[6] = v[6] * rstd * g1.z + b1.z; v[7] = v[7] * rstd * g1.w + b1.w;
        uint4 hi, lo;
        split_bf16x2(v[0], v[1], hi.x, lo.x); split_bf16x2(v[2], v[3], hi.y, lo.y);
        split_bf16x2(v[4], v[5], hi.z, lo.z); split_bf16x2(v[6], v[7], hi.w, lo.w);
        const int off = r * 128 + ((sub ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dA + off) = hi;
        *reinterpret_cast<uint4*>(dA + FF_PLANE + off) = lo;
      }
      ptx::mbar_arrive(&x_empty);               // staged rows consumed: the next tile may land
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&a_full[s]);
    }
  } else if (warp < FF_LOAD_WARPS + FF_EPI_WARPS) {
    // ================= epilogue warps =================
    const int ew = warp - FF_LOAD_WARPS;                // 0..15
    const int wq = warp & 3, cg = ew >> 2;              // TMEM lane quarter (hardware: warp % 4), column group
    const int row = wq * 32 + lane;
    const int cq = lane & 15;
    const float4 b2 = ldg4(a.b2 + cq * 4);
    float4 pg = make_float4(0, 0, 0, 0), pb = pg;
    if (a.pn_g) { pg = ldg4(a.pn_g + cq * 4); pb = ldg4(a.pn_b + cq * 4); }
    float4* stg = reinterpret_cast<float4*>(sH);        // [128 rows][16 x float4], chunk index XOR (row & 7)
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
      const int ab = it & 1;
      // ---- mid epilogues: acc1 -> +b1 -> swish -> H operand (K-chunk q of GEMM 2)
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int b = q & 1;
        const uint32_t u1 = (uint32_t)(2 * it + (q >> 1));
        const float* bias = a.b1 + q * 64 + cg * FF_CPT;
        float4 bb[FF_CPT / 4];
#pragma unroll
        for (int j = 0; j < FF_CPT / 4; ++j) bb[j] = ldg4(bias + 4 * j);
        ptx::mbar_wait(&acc1_full[b], u1 & 1u);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * 64 + cg * FF_CPT);
        float v[FF_CPT];
#pragma unroll
        for (int j = 0; j < FF_CPT; j += 8) {
          float t8[8];
          ptx::tmem_ld8(taddr + j, t8);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j + i] = t8[i];
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&acc1_empty[b]);
#pragma unroll
        for (int j = 0; j < FF_CPT / 4; ++j) { v[4 * j] += bb[j].x; v[4 * j + 1] += bb[j].y; v[4 * j + 2] += bb[j].z; v[4 * j + 3] += bb[j].w; }
#pragma unroll
        for (int j = 0; j < FF_CPT; ++j) v[j] *= sigmoidf_acc(v[j]);
        uint8_t* dH = sH;                                     // single H buffer; its use count is 4 * it + q
        const uint32_t uh = (uint32_t)(4 * it + q);
        ptx::mbar_wait(&h_empty, (uh & 1u) ^ 1u);
#pragma unroll
        for (int c8 = 0; c8 < FF_CPT / 8; ++c8) {
          uint4 hi, lo;
          split_bf16x2(v[c8 * 8 + 0], v[c8 * 8 + 1], hi.x, lo.x); split_bf16x2(v[c8 * 8 + 2], v[c8 * 8 + 3], hi.y, lo.y);
          split_bf16x2(v[c8 * 8 + 4], v[c8 * 8 + 5], hi.z, lo.z); split_bf16x2(v[c8 * 8 + 6], v[c8 * 8 + 7], hi.w, lo.w);
          const int c = cg * (FF_CPT / 8) + c8;
          const int off = row * 128 + ((c ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(dH + off) = hi;
          *reinterpret_cast<uint4*>(dH + FF_PLANE + off) = lo;
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&h_full);
      }
      // ---- final epilogue: acc2 -> staging (H buffer: every MMA2 of this tile has completed) -> coalesced store
      // issue the residual loads first: they do not depend on the accumulator
      float4 xv[128 / (FF_EPI_WARPS * 2)], r2v[128 / (FF_EPI_WARPS * 2)];
#pragma unroll
      for (int i8 = 0; i8 < 128 / (FF_EPI_WARPS * 2); ++i8) {
        const int m = m0 + i8 * (FF_EPI_WARPS * 2) + ew * 2 + (lane >> 4);
        const bool ok = m < a.M;
        xv[i8] = ok ? *reinterpret_cast<const float4*>(a.x + (long long)m * 64 + cq * 4) : make_float4(0, 0, 0, 0);
        r2v[i8] = (ok && a.pn_g) ? *reinterpret_cast<const float4*>(a.resid2 + (long long)m * 64 + cq * 4) : make_float4(0, 0, 0, 0);
      }
      ptx::mbar_wait(&acc2_full[ab], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(128 + ab * 64 + cg * FF_CPT);
#pragma unroll
        for (int j = 0; j < FF_CPT; j += 8) {
          float t8[8];
          ptx::tmem_ld8(taddr + j, t8);
          const int c0 = cg * (FF_CPT / 4) + (j >> 2);
          stg[row * 16 + ((c0 + 0) ^ (row & 7))] = make_float4(t8[0], t8[1], t8[2], t8[3]);
          stg[row * 16 + ((c0 + 1) ^ (row & 7))] = make_float4(t8[4], t8[5], t8[6], t8[7]);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc2_empty[ab]);
      asm volatile("bar.sync 1, %0;" ::"n"(FF_EPI_THREADS) : "memory");
#pragma unroll
      for (int i8 = 0; i8 < 128 / (FF_EPI_WARPS * 2); ++i8) {
        const int R = i8 * (FF_EPI_WARPS * 2) + ew * 2 + (lane >> 4);
        const int m = m0 + R;
        const bool ok = m < a.M;
        const float4 acc = stg[R * 16 + (cq ^ (R & 7))];
        float4 y;
        y.x = fmaf(a.alpha, acc.x + b2.x, xv[i8].x); y.y = fmaf(a.alpha, acc.y + b2.y, xv[i8].y);
        y.z = fmaf(a.alpha, acc.z + b2.z, xv[i8].z); y.w = fmaf(a.alpha, acc.w + b2.w, xv[i8].w);
        if (a.pn_g) {       // post_norm + outer residual (uniform branch)
          float sm = y.x + y.y + y.z + y.w;
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
          const float mean = sm * (1.0f / 64.0f);
          y.x -= mean; y.y -= mean; y.z -= mean; y.w -= mean;
          float qv = y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
          const float rstd = 1.0f / sqrtf(qv * (1.0f / 64.0f) + 1e-5f);
          y.x = y.x * rstd * pg.x + pb.x + r2v[i8].x; y.y = y.y * rstd * pg.y + pb.y + r2v[i8].y;
          y.z = y.z * rstd * pg.z + pb.z + r2v[i8].z; y.w = y.w * rstd * pg.w + pb.w + r2v[i8].w;
        }
        if (ok) st4(a.out + (long long)m * 64 + cq * 4, y);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(FF_EPI_THREADS) : "memory");     // staging tile is overwritten by the next tile's H chunk
    }
  } else if (warp == FF_LOAD_WARPS + FF_EPI_WARPS) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t uA = ptx::smem_u32(sA), uW1 = ptx::smem_u32(sW1), uH = ptx::smem_u32(sH), uW2 = ptx::smem_u32(sW2);
      auto gemm64 = [&](uint32_t d_tmem, uint32_t a_base, uint32_t w_base, bool fresh) {
        const uint64_t a_hi = ptx::umma_desc_sw128(a_base), a_lo = ptx::umma_desc_sw128(a_base + FF_PLANE);
        const uint64_t w_hi = ptx::umma_desc_sw128(w_base), w_lo = ptx::umma_desc_sw128(w_base + 64 * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ko = (uint64_t)((k * 32) >> 4);
          ptx::mma_bf16(d_tmem, a_lo + ko, w_hi + ko, IDESC, (fresh && k == 0) ? 0u : 1u);
          ptx::mma_bf16(d_tmem, a_hi + ko, w_lo + ko, IDESC, 1u);
          ptx::mma_bf16(d_tmem, a_hi + ko, w_hi + ko, IDESC, 1u);
        }
      };
      ptx::mbar_wait(&w_full, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1, ab = it & 1;
        auto mma1 = [&](int q) {
          const uint32_t u1 = (uint32_t)(2 * it + (q >> 1));
          ptx::mbar_wait(&acc1_empty[q & 1], (u1 & 1u) ^ 1u);
          ptx::tc_fence_after();
          gemm64(tmem_base + (uint32_t)((q & 1) * 64), uA + s * 2 * FF_PLANE, uW1 + q * FF_WBLK, true);
          ptx::tc_commit(&acc1_full[q & 1]);
          if (q == 3) ptx::tc_commit(&a_empty[s]);
        };
        auto mma2 = [&](int q) {
          const uint32_t u1 = (uint32_t)(2 * it + (q >> 1));      // use count of W2 ring slot q & 1
          ptx::mbar_wait(&h_full, (uint32_t)(4 * it + q) & 1u);
          ptx::mbar_wait(&w2_full[q & 1], u1 & 1u);
          ptx::tc_fence_after();
          gemm64(tmem_base + 128u + (uint32_t)(ab * 64), uH, uW2 + (q & 1) * FF_WBLK, q == 0);
          ptx::tc_commit(&h_empty);
          ptx::tc_commit(&w2_empty[q & 1]);
          if (q == 3) ptx::tc_commit(&acc2_full[ab]);
        };
        ptx::mbar_wait(&a_full[s], (uint32_t)(it >> 1) & 1u);
        ptx::mbar_wait(&acc2_empty[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        mma1(0);
        for (int q = 0; q < 4; ++q) {
          if (q < 3) mma1(q + 1);
          mma2(q);
        }
      }
    }
  } else if (warp == FF_LOAD_WARPS + FF_EPI_WARPS + 2) {
    // ================= x staging: one bulk copy per tile (rows are contiguous in HBM) =================
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it) {
        const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
        const int rows = (a.M - m0 < BM) ? a.M - m0 : BM;
        ptx::mbar_wait(&x_empty, ((uint32_t)it & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&x_full, (uint32_t)rows * 256u);
        ptx::bulk_g2s(ptx::smem_u32(sX), a.x + (long long)m0 * 64, (uint32_t)rows * 256u, &x_full);
      }
    }
  } else {
    // ================= weights: loaded once, resident for the whole kernel =================
    if (lane == 0 && my_tiles > 0) {
      ptx::mbar_arrive_expect_tx(&w_full, 4 * FF_WBLK);           // W1: loaded once, resident
      for (int q = 0; q < 4; ++q) ptx::bulk_g2s(ptx::smem_u32(sW1) + q * FF_WBLK, a.w1 + (size_t)q * FF_WBLK, FF_WBLK, &w_full);
      for (int it = 0; it < my_tiles; ++it) {                     // W2: block q of every tile through the 2-slot ring
        for (int q = 0; q < 4; ++q) {
          const uint32_t u = (uint32_t)(2 * it + (q >> 1));
          ptx::mbar_wait(&w2_empty[q & 1], (u & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&w2_full[q & 1], FF_WBLK);
          ptx::bulk_g2s(ptx::smem_u32(sW2) + (q & 1) * FF_WBLK, a.w2 + (size_t)q * FF_WBLK, FF_WBLK, &w2_full[q & 1]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == FF_LOAD_WARPS + FF_EPI_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_ffn_fused(const SebFfn* f, void* stream) {
  SEB_REQUIRE(f && f->x && f->out && f->tokens > 0 && f->tokens < 2147483647LL - 128, SEB_EINVAL, "ffn_fused: bad arguments");
  SEB_REQUIRE(f->ln_gamma && f->ln_beta && f->w1_tc && f->b1 && f->w2_tc && f->b2, SEB_EINVAL, "ffn_fused: null parameter");
  SEB_REQUIRE(aligned16(f->x) && aligned16(f->out) && aligned16(f->w1_tc) && aligned16(f->w2_tc) && aligned16(f->b1) && aligned16(f->b2), SEB_EALIGN, "ffn_fused: unaligned pointer");
  if (f->post_gamma) SEB_REQUIRE(f->post_beta && f->resid2 && aligned16(f->resid2), SEB_EINVAL, "ffn_fused: post-norm needs beta and resid2");
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM);
    if (e != cudaSuccess) { set_error("ffn_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
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
  ffn_fused_kernel<<<grid, FF_THREADS, FF_SMEM, (cudaStream_t)stream>>>(a);
  SEB_CHECK_LAUNCH("ffn_fused_kernel");
  return 0;
}
