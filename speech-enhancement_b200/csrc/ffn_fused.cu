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
// acc1 is double buffered, so MMA1(q+1) overlaps the Swish epilogue of chunk q; weights stream through 16 KB
// single-slot rings with cp.async.bulk + mbarriers.  96 KB smem and 256 TMEM columns per CTA -> 2 CTAs per SM.
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

constexpr int FF_THREADS = 320;
constexpr int FF_PLANE = BM * 128;                 // 16 KB: one bf16 plane of a 128 x 64 operand tile
constexpr int FF_WBLK = 2 * 64 * 128;              // 16 KB: hi|lo image of a 64-row x 64-k weight block
constexpr int FF_SMEM = 1024 + 2 * FF_PLANE + FF_WBLK + 2 * FF_PLANE + FF_WBLK;

__global__ void __launch_bounds__(FF_THREADS, 2) ffn_fused_kernel(const FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full, w1_full, w1_empty, w2_full, w2_empty, h_full, h_empty, acc1_full[2], acc1_empty[2], acc2_full;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // A hi | lo   (32 KB)   -- reused as the fp32 staging tile at the end
  uint8_t* sW1 = sA + 2 * FF_PLANE;         // W1 block hi | lo (16 KB)
  uint8_t* sH = sW1 + FF_WBLK;              // H hi | lo   (32 KB)
  uint8_t* sW2 = sH + 2 * FF_PLANE;         // W2 block hi | lo (16 KB)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;

  if (tid == 0) {
    ptx::mbar_init(&a_full, 256);
    ptx::mbar_init(&w1_full, 1); ptx::mbar_init(&w1_empty, 1);
    ptx::mbar_init(&w2_full, 1); ptx::mbar_init(&w2_empty, 1);
    ptx::mbar_init(&h_full, 256); ptx::mbar_init(&h_empty, 1);
    ptx::mbar_init(&acc1_full[0], 1); ptx::mbar_init(&acc1_full[1], 1);
    ptx::mbar_init(&acc1_empty[0], 256); ptx::mbar_init(&acc1_empty[1], 256);
    ptx::mbar_init(&acc2_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc(&tmem_base_s, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ---------------- A operand: LayerNorm(x) split to bf16 hi/lo ----------------
    {
      GemmArgs g;
      g.a[0] = a.x; g.lda = 64; g.M = a.M; g.ln_g = a.ln_g; g.ln_b = a.ln_b;
      const int sub = tid & 7, rloc = tid >> 3;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        Loader<SEB_LOAD_ROWS_LN>::Row row;
        Loader<SEB_LOAD_ROWS_LN>::init_row(g, m0 + p * 32 + rloc, row);
        float v[8];
        Loader<SEB_LOAD_ROWS_LN>::load(g, row, 0, sub, v);
        uint4 hi, lo;
        split_bf16x2(v[0], v[1], hi.x, lo.x); split_bf16x2(v[2], v[3], hi.y, lo.y);
        split_bf16x2(v[4], v[5], hi.z, lo.z); split_bf16x2(v[6], v[7], hi.w, lo.w);
        const int r = p * 32 + rloc;
        const int off = r * 128 + ((sub ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sA + off) = hi;
        *reinterpret_cast<uint4*>(sA + FF_PLANE + off) = lo;
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&a_full);
    }
    // ---------------- mid epilogue: acc1 -> +b1 -> swish -> H operand ----------------
    const int wq = warp & 3, half = warp >> 2;
    const int row = wq * 32 + lane;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int b = q & 1;
      ptx::mbar_wait(&acc1_full[b], (uint32_t)(q >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * 64 + half * 32);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t8[8];
        ptx::tmem_ld8(taddr + j, t8);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j + i] = t8[i];
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc1_empty[b]);
      const float* bias = a.b1 + q * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = ldg4(bias + j);
        v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= sigmoidf_acc(v[j]);
      ptx::mbar_wait(&h_empty, (uint32_t)(q & 1) ^ 1u);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint4 hi, lo;
        split_bf16x2(v[c8 * 8 + 0], v[c8 * 8 + 1], hi.x, lo.x); split_bf16x2(v[c8 * 8 + 2], v[c8 * 8 + 3], hi.y, lo.y);
        split_bf16x2(v[c8 * 8 + 4], v[c8 * 8 + 5], hi.z, lo.z); split_bf16x2(v[c8 * 8 + 6], v[c8 * 8 + 7], hi.w, lo.w);
        const int c = half * 4 + c8;
        const int off = row * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(sH + off) = hi;
        *reinterpret_cast<uint4*>(sH + FF_PLANE + off) = lo;
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&h_full);
    }
    // ---------------- final epilogue ----------------
    ptx::mbar_wait(&acc2_full, 0);
    ptx::tc_fence_after();
    float4* stg = reinterpret_cast<float4*>(sA);           // [128 rows][16 x float4], chunk index XOR (row & 7)
    {
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(128 + half * 32);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t8[8];
        ptx::tmem_ld8(taddr + j, t8);
        const int c0 = half * 8 + (j >> 2);
        stg[row * 16 + ((c0 + 0) ^ (row & 7))] = make_float4(t8[0], t8[1], t8[2], t8[3]);
        stg[row * 16 + ((c0 + 1) ^ (row & 7))] = make_float4(t8[4], t8[5], t8[6], t8[7]);
      }
    }
    ptx::tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int cq = lane & 15;
    const float4 b2 = ldg4(a.b2 + cq * 4);
    float4 pg = make_float4(0, 0, 0, 0), pb = pg;
    if (a.pn_g) { pg = ldg4(a.pn_g + cq * 4); pb = ldg4(a.pn_b + cq * 4); }
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
      const int R = it * 16 + warp * 2 + (lane >> 4);
      const int m = m0 + R;
      const bool ok = m < a.M;
      float4 acc = stg[R * 16 + (cq ^ (R & 7))];
      float4 xv = ok ? *reinterpret_cast<const float4*>(a.x + (long long)m * 64 + cq * 4) : make_float4(0, 0, 0, 0);
      float4 y;
      y.x = fmaf(a.alpha, acc.x + b2.x, xv.x); y.y = fmaf(a.alpha, acc.y + b2.y, xv.y);
      y.z = fmaf(a.alpha, acc.z + b2.z, xv.z); y.w = fmaf(a.alpha, acc.w + b2.w, xv.w);
      if (a.pn_g) {       // post_norm + outer residual (uniform branch)
        float s = y.x + y.y + y.z + y.w;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 64.0f);
        y.x -= mean; y.y -= mean; y.z -= mean; y.w -= mean;
        float qv = y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
        const float rstd = 1.0f / sqrtf(qv * (1.0f / 64.0f) + 1e-5f);
        float4 r2 = ok ? *reinterpret_cast<const float4*>(a.resid2 + (long long)m * 64 + cq * 4) : make_float4(0, 0, 0, 0);
        y.x = y.x * rstd * pg.x + pb.x + r2.x; y.y = y.y * rstd * pg.y + pb.y + r2.y;
        y.z = y.z * rstd * pg.z + pb.z + r2.z; y.w = y.w * rstd * pg.w + pb.w + r2.w;
      }
      if (ok) st4(a.out + (long long)m * 64 + cq * 4, y);
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
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
      auto mma1 = [&](int q) {
        ptx::mbar_wait(&w1_full, (uint32_t)q & 1u);
        ptx::mbar_wait(&acc1_empty[q & 1], ((uint32_t)(q >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        gemm64(tmem_base + (uint32_t)((q & 1) * 64), uA, uW1, true);
        ptx::tc_commit(&w1_empty);
        ptx::tc_commit(&acc1_full[q & 1]);
      };
      auto mma2 = [&](int q) {
        ptx::mbar_wait(&h_full, (uint32_t)q & 1u);
        ptx::mbar_wait(&w2_full, (uint32_t)q & 1u);
        ptx::tc_fence_after();
        gemm64(tmem_base + 128u, uH, uW2, q == 0);
        ptx::tc_commit(&h_empty);
        ptx::tc_commit(&w2_empty);
        if (q == 3) ptx::tc_commit(&acc2_full);
      };
      ptx::mbar_wait(&a_full, 0);
      ptx::tc_fence_after();
      mma1(0);
      for (int q = 0; q < 4; ++q) {
        if (q < 3) mma1(q + 1);
        mma2(q);
      }
    }
  } else {
    // ---------------- weight stager ----------------
    if (lane == 0) {
      const uint32_t uW1 = ptx::smem_u32(sW1), uW2 = ptx::smem_u32(sW2);
      for (int q = 0; q < 4; ++q) {
        ptx::mbar_wait(&w1_empty, ((uint32_t)q & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&w1_full, FF_WBLK);
        ptx::bulk_g2s(uW1, a.w1 + (size_t)q * FF_WBLK, FF_WBLK, &w1_full);
        ptx::mbar_wait(&w2_empty, ((uint32_t)q & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&w2_full, FF_WBLK);
        ptx::bulk_g2s(uW2, a.w2 + (size_t)q * FF_WBLK, FF_WBLK, &w2_full);
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
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
  const unsigned grid = (unsigned)((f->tokens + BM - 1) / BM);
  ffn_fused_kernel<<<grid, FF_THREADS, FF_SMEM, (cudaStream_t)stream>>>(a);
  SEB_CHECK_LAUNCH("ffn_fused_kernel");
  return 0;
}
