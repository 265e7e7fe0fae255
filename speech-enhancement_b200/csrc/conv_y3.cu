// conv_y3.cu -- dilated dense convolution (generator.py:16-20, stride 1) with the three frequency taps folded into
// the N dimension of ONE tcgen05.mma.
//
// conv_persist / conv_split run three N = 64 MMAs per K-step, one per frequency tap, each with its own copy of the
// activations: 32 KB of A + 16 KB of W per 384 MMA-cycles, which saturates the L2 -> SM path (measured 10.3 TB/s, tensor
// pipe 33 %).  Here the A tile (128 consecutive pixels x 64 channels) is loaded ONCE per (time-tap, slot) stage and
// multiplied by the stacked weights [W_kf=0; W_kf=1; W_kf=2] (N = 192):
//       Y_kf[q] = W_kf . x[q]                      (three 64-column blocks of one 128 x 192 accumulator)
//       out[p]  = Y_0[p - 1] + Y_1[p] + Y_2[p + 1] (frequency taps f-1, f, f+1 are the flat pixels p-1, p, p+1)
// The +-1 row shifts are applied in the epilogue on the fp32 staging tile, with the image-row borders masked there
// (f == 0 has no left tap, f == F-1 no right tap).  A tile of 128 input pixels yields 126 outputs.  Per stage:
// 32 KB A + 48 KB W for 1152 MMA-cycles = 1.8x less L2 traffic per MMA-cycle.  Persistent, warp-specialised:
// 2-slot ring (80 KB stages), 2 producer groups x 4 warps (cp.async), weight thread (cp.async.bulk), MMA thread,
// 8 epilogue warps, double-buffered 192-column TMEM accumulator.
#include "gemm_engine.cuh"

namespace seb {

constexpr int Y3_SLOTS = 2;
constexpr int Y3_GW = 4;
constexpr int Y3_PROD_WARPS = Y3_SLOTS * Y3_GW;            // 8
constexpr int Y3_EPI_WARPS = 8;
constexpr int Y3_EPI_THREADS = Y3_EPI_WARPS * 32;
constexpr int Y3_THREADS = (Y3_PROD_WARPS + Y3_EPI_WARPS + 2) * 32;      // 576
constexpr int Y3_WPLANE = 192 * 128;                                      // 24 KB: one bf16 plane of the stacked weights
constexpr int Y3_STAGE = 2 * TC_A_BYTES + 2 * Y3_WPLANE;                  // 80 KB
constexpr int Y3_SMEM = Y3_SLOTS * Y3_STAGE + BM * 64 * 4 + 1024;         // ring + fp32 staging (one 64-column block) = 197,632 B
constexpr int Y3_OUT = BM - 2;                                            // 126 outputs per tile
constexpr int Y3_RPP = Y3_GW * 2, Y3_NPASS = BM / Y3_RPP;                 // 8 rows per pass, 16 passes

__global__ void __launch_bounds__(Y3_THREADS, 1) conv_y3_kernel(const GemmArgs g, const uint8_t* __restrict__ w_tc) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[Y3_SLOTS], empty_bar[Y3_SLOTS], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS/LDS, not generic ST/LD)
  const uint32_t smem_base = ptx::smem_u32(smem);
  float4* stg = reinterpret_cast<float4*>(smem + Y3_SLOTS * Y3_STAGE);    // [128 rows][16 x float4], chunk XOR (row & 7)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = g.Fin;
  const int nst = g.taps_t * g.nslots;            // stages per tile: (kt, slot)
  const int ntiles = (g.M + Y3_OUT - 1) / Y3_OUT;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < Y3_SLOTS; ++s) { ptx::mbar_init(&full_bar[s], Y3_GW + 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], Y3_EPI_THREADS); }
    ptx::fence_barrier_init();
  }
  if (warp == Y3_PROD_WARPS + Y3_EPI_WARPS) ptx::tmem_alloc(&tmem_base_s, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;         // accumulator ab at column 256 * ab (192 used)

  if (warp < Y3_PROD_WARPS) {
    // ================= producer group gidx: ring slot gidx, global stages gidx, gidx + 2, ... =================
    const int gidx = warp / Y3_GW, gw = warp % Y3_GW;
    const int c = lane & 7, plane = (lane >> 3) & 1, r0 = gw * 2 + (lane >> 4);
    const uint32_t dst0 = smem_base + gidx * Y3_STAGE + (uint32_t)(plane * TC_A_BYTES + r0 * 128 + ((c ^ (r0 & 7)) << 4));
    const int src_lane_off = plane * 128 + c * 16;
    int pix[Y3_NPASS], trow[Y3_NPASS];            // input pixel of row r (-1: outside the tensor) and its t
    int cur_tile = -1;
    const long long total = (long long)my_tiles * nst;
    for (long long gs = gidx; gs < total; gs += Y3_SLOTS) {
      const int it = (int)(gs / nst), st = (int)(gs - (long long)it * nst);
      if (it != cur_tile) {
        cur_tile = it;
        const int q0 = ((int)blockIdx.x + it * (int)gridDim.x) * Y3_OUT - 1 + r0;      // first row of this thread (may be -1)
        const int qs = q0 + F;                                                        // shifted by one image row: non-negative
        int bt = qs / F - 1, f = qs - (bt + 1) * F;
        int t = (bt + g.T) % g.T;
#pragma unroll
        for (int p = 0; p < Y3_NPASS; ++p) {
          const int q = q0 + p * Y3_RPP;
          pix[p] = (q >= 0 && q < g.M) ? q : -1;
          trow[p] = t;
          f += Y3_RPP;
          if (f >= F) { f -= F; if (++t == g.T) t = 0; }
        }
      }
      const uint32_t ph = (uint32_t)(gs / Y3_SLOTS) & 1u;
      const int kt = st / g.nslots, slot = st - kt * g.nslots;
      const int dt = (g.taps_t - 1 - kt) * g.dil;
      const uint8_t* src_base = reinterpret_cast<const uint8_t*>(g.a[slot]) + src_lane_off;
      const int dpix = -dt * F;
      ptx::mbar_wait(&empty_bar[gidx], ph ^ 1u);
#pragma unroll
      for (int p = 0; p < Y3_NPASS; ++p) {
        const bool ok = pix[p] >= 0 && trow[p] >= dt;
        const long long q = ok ? (long long)(pix[p] + dpix) : 0;
        cp_async16_zfill(dst0 + p * (Y3_RPP * 128), src_base + q * 256, ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_bar[gidx]);
    }
  } else if (warp < Y3_PROD_WARPS + Y3_EPI_WARPS) {
    // ================= epilogue: out[p] = Y0[p-1] + Y1[p] + Y2[p+1] (+ bias), three passes over the staging tile =================
    const int ew = warp - Y3_PROD_WARPS;              // 0..7
    const int wq = warp & 3, half = ew >> 2;          // TMEM lane quarter (hardware: warp % 4; Y3_PROD_WARPS % 4 == 0), column half
    const int row = wq * 32 + lane;
    const int cq = lane & 15;
    float4 bias = make_float4(0, 0, 0, 0);
    if (g.bias) bias = ldg4(g.bias + cq * 4);
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * Y3_OUT;      // first output pixel; staging row i+kf holds input pixel m0 - 1 + i + kf
      const int ab = it & 1;
      ptx::mbar_wait(&acc_full[ab], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      float4 acc[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) acc[i8] = bias;
      // f of this thread's 8 output pixels (rows i = i8*16 + ew*2 + (lane >> 4))
      int fo[8];
      {
        const int i0 = ew * 2 + (lane >> 4);
        int f = (m0 + i0) % F;
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) { fo[i8] = f; f += 16; if (f >= F) f -= F; }
      }
#pragma unroll 1
      for (int kf = 0; kf < 3; ++kf) {
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * 256 + kf * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8];
          ptx::tmem_ld8(taddr + j, v);
          const int c0 = half * 8 + (j >> 2);
          stg[row * 16 + ((c0 + 0) ^ (row & 7))] = make_float4(v[0], v[1], v[2], v[3]);
          stg[row * 16 + ((c0 + 1) ^ (row & 7))] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (kf == 2) { ptx::tc_fence_before(); ptx::mbar_arrive(&acc_empty[ab]); }
        asm volatile("bar.sync 1, %0;" ::"n"(Y3_EPI_THREADS) : "memory");
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int i = i8 * 16 + ew * 2 + (lane >> 4);            // output index inside the tile
          const int R = i + kf;                                    // staging row: Y_kf of input pixel p - 1 + kf
          const bool ok = (i < Y3_OUT) && (kf == 1 || (kf == 0 ? fo[i8] >= 1 : fo[i8] <= F - 2));
          if (ok) {
            const float4 v = stg[R * 16 + (cq ^ (R & 7))];
            acc[i8].x += v.x; acc[i8].y += v.y; acc[i8].z += v.z; acc[i8].w += v.w;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(Y3_EPI_THREADS) : "memory");
      }
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int i = i8 * 16 + ew * 2 + (lane >> 4);
        const int p = m0 + i;
        if (i < Y3_OUT && p < g.M) st4(g.out + (long long)p * g.ldo + cq * 4, acc[i8]);
      }
    }
  } else if (warp == Y3_PROD_WARPS + Y3_EPI_WARPS) {
    // ================= MMA issuer: M = 128, N = 192 (three taps), K = 16 =================
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      long long gs = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int ab = it & 1;
        ptx::mbar_wait(&acc_empty[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 256);
        for (int st = 0; st < nst; ++st, ++gs) {
          const int s = (int)(gs % Y3_SLOTS);
          const uint32_t ph = (uint32_t)(gs / Y3_SLOTS) & 1u;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t base = smem_base + s * Y3_STAGE;
          const uint64_t a_hi = ptx::umma_desc_sw128(base), a_lo = ptx::umma_desc_sw128(base + TC_A_BYTES);
          const uint64_t w_hi = ptx::umma_desc_sw128(base + 2 * TC_A_BYTES), w_lo = ptx::umma_desc_sw128(base + 2 * TC_A_BYTES + Y3_WPLANE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ptx::mma_bf16(d_tmem, a_lo + ko, w_hi + ko, IDESC, (st | k) ? 1u : 0u);
            ptx::mma_bf16(d_tmem, a_hi + ko, w_lo + ko, IDESC, 1u);
            ptx::mma_bf16(d_tmem, a_hi + ko, w_hi + ko, IDESC, 1u);
          }
          ptx::tc_commit(&empty_bar[s]);
        }
        ptx::tc_commit(&acc_full[ab]);
      }
    }
  } else {
    // ================= weight stager: stack the three tap blocks of stage (kt, slot): hi plane | lo plane =================
    if (lane == 0) {
      long long gs = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int st = 0; st < nst; ++st, ++gs) {
          const int s = (int)(gs % Y3_SLOTS);
          const uint32_t ph = (uint32_t)(gs / Y3_SLOTS) & 1u;
          const int kt = st / g.nslots, slot = st - kt * g.nslots;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * Y3_WPLANE);
          const uint32_t wb = smem_base + s * Y3_STAGE + 2 * TC_A_BYTES;
#pragma unroll
          for (int kf = 0; kf < 3; ++kf) {
            const int kc = (kt * 3 + kf) * g.nslots + slot;        // K order of the packed image: (kt, kf, slot); block = hi 8 KB | lo 8 KB
            const uint8_t* src = w_tc + (size_t)kc * (2 * 64 * 128);
            ptx::bulk_g2s(wb + kf * (64 * 128), src, 64 * 128, &full_bar[s]);
            ptx::bulk_g2s(wb + Y3_WPLANE + kf * (64 * 128), src + 64 * 128, 64 * 128, &full_bar[s]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == Y3_PROD_WARPS + Y3_EPI_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

int launch_conv_y3(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  static int num_sms = 0;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(conv_y3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Y3_SMEM);
    if (e != cudaSuccess) { set_error("conv y3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    attr_done.set();
  }
  SEB_REQUIRE(s->Fin >= 16 && (long long)s->M + s->Fin < 2147483647LL, SEB_EINVAL, "conv y3: geometry out of range");
  const long long ntiles = ((long long)s->M + Y3_OUT - 1) / Y3_OUT;
  dim3 grid((unsigned)(ntiles < num_sms ? ntiles : num_sms));
  conv_y3_kernel<<<grid, Y3_THREADS, Y3_SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc));
  SEB_CHECK_LAUNCH("conv_y3_kernel");
  return 0;
}

}  // namespace seb
