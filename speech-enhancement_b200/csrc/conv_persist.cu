// conv_persist.cu -- persistent implicit-GEMM convolution on pre-split activations (dilated dense convs and conv_2).
//
// Measured on B200 (tools/conv_probe.py, profiles/r1): the one-tile-per-CTA kernel (conv_split_tc_kernel) spends ~35 %
// of its time in per-tile fixed cost (launch, barrier init, TMEM alloc, index set-up, un-overlapped epilogue) and its
// two ring slots per CTA cannot cover the ~3-4.5 k-cycle producer <-> MMA hand-off round trip of a 384-cycle chunk.
// This kernel keeps one CTA per SM alive over all its tiles:
//   * a 4-slot ring of (A hi|lo 32 KB + W hi|lo 16 KB) stages; slot g is owned by producer group g (4 warps), which
//     loads every 4th K-chunk -- publishing a chunk depends only on MMA(c - 4) and on its own cp.async copies;
//   * the producers run straight on into the next tile (the ring never drains between tiles);
//   * one thread streams weight blocks (cp.async.bulk), one thread issues the tcgen05.mma triples;
//   * 4 epilogue warps drain a double-buffered TMEM accumulator (tile i is stored while tile i+1 is multiplied).
#include "gemm_engine.cuh"

namespace seb {

constexpr int CP_SLOTS = 4;
constexpr int CP_GW = 4;                                   // warps per producer group
constexpr int CP_PROD_WARPS = CP_SLOTS * CP_GW;            // 16
constexpr int CP_EPI_WARPS = 4;
constexpr int CP_THREADS = (CP_PROD_WARPS + CP_EPI_WARPS + 2) * 32;      // 704
constexpr int CP_STAGE = 2 * TC_A_BYTES + 2 * 64 * 128;                  // 48 KB
constexpr int CP_SMEM = CP_SLOTS * CP_STAGE + BM * 64 * 4 + 1024;        // ring + fp32 epilogue staging = 230,400 B
constexpr int CP_RPP = CP_GW * 2, CP_NPASS = BM / CP_RPP;                // 8 rows per pass, 16 passes

__global__ void __launch_bounds__(CP_THREADS, 1) conv_persist_kernel(const GemmArgs g, const uint8_t* __restrict__ w_tc) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[CP_SLOTS], empty_bar[CP_SLOTS], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS/LDS, not generic ST/LD)
  const uint32_t smem_base = ptx::smem_u32(smem);
  uint8_t* stg_base = smem + CP_SLOTS * CP_STAGE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkc = g.K / BK;                       // K-chunks per tile
  const int ntiles = (g.M + BM - 1) / BM;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < CP_SLOTS; ++s) { ptx::mbar_init(&full_bar[s], CP_GW + 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], CP_EPI_WARPS * 32); }
    ptx::fence_barrier_init();
  }
  if (warp == CP_PROD_WARPS + CP_EPI_WARPS) ptx::tmem_alloc(&tmem_base_s, 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;         // accumulator ab at column 64 * ab

  if (warp < CP_PROD_WARPS) {
    // ================= producer group gidx: ring slot gidx, global chunks gidx, gidx + 4, ... =================
    const int gidx = warp / CP_GW, gw = warp % CP_GW;
    const int c = lane & 7, plane = (lane >> 3) & 1, r0 = gw * 2 + (lane >> 4);
    const uint32_t dst0 = smem_base + gidx * CP_STAGE + (uint32_t)(plane * TC_A_BYTES + r0 * 128 + ((c ^ (r0 & 7)) << 4));
    const int src_lane_off = plane * 128 + c * 16;
    int pix[CP_NPASS], tf[CP_NPASS];
    int cur_tile = -1;
    const long long total = (long long)my_tiles * nkc;
    for (long long gc = gidx; gc < total; gc += CP_SLOTS) {
      const int it = (int)(gc / nkc), kc = (int)(gc - (long long)it * nkc);
      if (it != cur_tile) {                       // row table of this thread's 16 rows (no per-row divisions)
        cur_tile = it;
        const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
        const int bt0 = m0 / g.Fout, f0 = m0 - bt0 * g.Fout, t0 = bt0 % g.T;
#pragma unroll
        for (int p = 0; p < CP_NPASS; ++p) {
          const int r = p * CP_RPP + r0;
          int f = f0 + r, bt = bt0, t = t0;
          if (f >= g.Fout) { f -= g.Fout; ++bt; if (++t == g.T) t = 0; }
          if (f >= g.Fout) { f -= g.Fout; ++bt; if (++t == g.T) t = 0; }
          if (m0 + r < g.M) { pix[p] = bt * g.Fin + f * g.stride_f; tf[p] = (t << 16) | f; }
          else { pix[p] = -1; tf[p] = 0; }
        }
      }
      const uint32_t ph = (uint32_t)(gc / CP_SLOTS) & 1u;
      const int tap = kc / g.nslots, slot = kc - tap * g.nslots;
      const int kt = (g.taps_t == 2) ? tap / 3 : 0, kf = tap - kt * 3;
      const int dt = (g.taps_t - 1 - kt) * g.dil, df = kf - 1;
      const uint8_t* src_base = reinterpret_cast<const uint8_t*>(g.a[slot]) + src_lane_off;
      const int dpix = df - dt * g.Fin;
      ptx::mbar_wait(&empty_bar[gidx], ph ^ 1u);
#pragma unroll
      for (int p = 0; p < CP_NPASS; ++p) {
        const int t = tf[p] >> 16, f = tf[p] & 0xffff;
        const int ff = f * g.stride_f + df;
        const bool ok = pix[p] >= 0 && t >= dt && ff >= 0 && ff < g.Fin;
        const long long q = ok ? (long long)(pix[p] + dpix) : 0;
        cp_async16_zfill(dst0 + p * (CP_RPP * 128), src_base + q * 256, ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_bar[gidx]);
    }
  } else if (warp < CP_PROD_WARPS + CP_EPI_WARPS) {
    // ================= epilogue warps: TMEM -> smem transpose -> coalesced bias store =================
    const int wq = warp & 3;                      // TMEM lane quarter (hardware: warp % 4); CP_PROD_WARPS % 4 == 0
    float4* stg = reinterpret_cast<float4*>(stg_base + wq * 8192);      // [32 rows][16 x float4]
    const int cq = lane & 15;
    float4 bias = make_float4(0, 0, 0, 0);
    if (g.bias) bias = ldg4(g.bias + cq * 4);
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
      const int ab = it & 1;
      ptx::mbar_wait(&acc_full[ab], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * 64);
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        float v[8];
        ptx::tmem_ld8(taddr + j, v);
        stg[lane * 16 + (((j >> 2) + 0) ^ (lane & 7))] = make_float4(v[0], v[1], v[2], v[3]);
        stg[lane * 16 + (((j >> 2) + 1) ^ (lane & 7))] = make_float4(v[4], v[5], v[6], v[7]);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_empty[ab]);
      __syncwarp();
#pragma unroll 4
      for (int i2 = 0; i2 < 16; ++i2) {
        const int R = i2 * 2 + (lane >> 4);
        const int m = m0 + wq * 32 + R;
        if (m < g.M) {
          float4 val = stg[R * 16 + (cq ^ (R & 7))];
          val.x += bias.x; val.y += bias.y; val.z += bias.z; val.w += bias.w;
          st4(g.out + (long long)m * g.ldo + cq * 4, val);
        }
      }
      __syncwarp();
    }
  } else if (warp == CP_PROD_WARPS + CP_EPI_WARPS) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      long long gc = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int ab = it & 1;
        ptx::mbar_wait(&acc_empty[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 64);
        for (int kc = 0; kc < nkc; ++kc, ++gc) {
          const int s = (int)(gc % CP_SLOTS);
          const uint32_t ph = (uint32_t)(gc / CP_SLOTS) & 1u;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t base = smem_base + s * CP_STAGE;
          const uint64_t a_hi = ptx::umma_desc_sw128(base), a_lo = ptx::umma_desc_sw128(base + TC_A_BYTES);
          const uint64_t w_hi = ptx::umma_desc_sw128(base + 2 * TC_A_BYTES), w_lo = ptx::umma_desc_sw128(base + 2 * TC_A_BYTES + 64 * 128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ptx::mma_bf16(d_tmem, a_lo + ko, w_hi + ko, IDESC, (kc | k) ? 1u : 0u);
            ptx::mma_bf16(d_tmem, a_hi + ko, w_lo + ko, IDESC, 1u);
            ptx::mma_bf16(d_tmem, a_hi + ko, w_hi + ko, IDESC, 1u);
          }
          ptx::tc_commit(&empty_bar[s]);
        }
        ptx::tc_commit(&acc_full[ab]);
      }
    }
  } else {
    // ================= weight stager =================
    if (lane == 0) {
      long long gc = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int kc = 0; kc < nkc; ++kc, ++gc) {
          const int s = (int)(gc % CP_SLOTS);
          const uint32_t ph = (uint32_t)(gc / CP_SLOTS) & 1u;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * 64 * 128);
          ptx::bulk_g2s(smem_base + s * CP_STAGE + 2 * TC_A_BYTES, w_tc + (size_t)kc * (2 * 64 * 128), 2 * 64 * 128, &full_bar[s]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == CP_PROD_WARPS + CP_EPI_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

int launch_conv_persist(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  static int num_sms = 0;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(conv_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM);
    if (e != cudaSuccess) { set_error("conv persist: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    attr_done.set();
  }
  SEB_REQUIRE(s->T < 32768 && s->Fout < 65536 && s->Fout >= 64 && (long long)s->B * s->T * s->Fin < 2147483647LL, SEB_EINVAL,
              "conv persist: geometry outside the packed row-table range (needs 64 <= Fout < 65536, T < 32768)");
  const long long ntiles = ((long long)s->M + BM - 1) / BM;
  dim3 grid((unsigned)(ntiles < num_sms ? ntiles : num_sms));
  conv_persist_kernel<<<grid, CP_THREADS, CP_SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc));
  SEB_CHECK_LAUNCH("conv_persist_kernel");
  return 0;
}

}  // namespace seb
