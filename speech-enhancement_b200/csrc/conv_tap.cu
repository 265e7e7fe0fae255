// conv_tap.cu -- dilated dense convolution (generator.py:16-20) as an implicit GEMM whose A tile is shared by the three
// frequency taps.
//
// conv_split_tc_kernel re-reads the activations once per tap: 32 KB of A + 16 KB of W per 384 MMA-cycles, ~3x what the
// L2 -> SM path delivers (measured: tensor pipe 25-33 %).  Here one (kt, slot) stage stages the 128 output pixels PLUS
// ONE HALO PIXEL EACH SIDE once, and the taps kf = 0,1,2 are three tcgen05.mma groups whose A descriptors start 0, 16,
// 32 bytes into the tile: the A tile uses the UMMA K-major *no-swizzle* canonical layout with the 8-row core-matrix
// groups packed back to back (SBO = 128 B), so consecutive pixels are exactly 16 bytes apart inside every 8-channel
// K-chunk (chunks are LBO apart) and "shift by one pixel" is "start 16 bytes later".
//
// Image-row borders are handled by walking a PADDED pixel index p' = (b*T + t)*(F+1) + f: the virtual pixel f == F is
// zero-filled and serves as the right halo of row t and the left halo of row t+1; its outputs are never stored
// (1/(F+1) of the tile rows).  The time tap (t - dil) is a different stage with its own tile (zero-filled for t < dil).
#include "gemm_engine.cuh"
#include <stdlib.h>

namespace seb {

constexpr int CT_ROWS = BM + 2;                  // 130 staged pixels
constexpr int CT_LBO = 137 * 16;                 // bytes between the 8-channel K-chunks of the A tile (137 rows: 548 words = 4 mod 32 banks)
constexpr int CT_APLANE = 8 * CT_LBO;            // 17,536 B per bf16 plane
constexpr int CT_WTAP = 2 * 64 * 128;            // 16 KB: hi|lo weight block of one tap (swizzled, as packed)
constexpr int CT_STAGE = ((3 * CT_WTAP + 2 * CT_APLANE + 1023) / 1024) * 1024;   // 84,992 B

// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; 8-row groups SBO apart; the two K-chunks of one MMA LBO apart
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Persistent, warp-specialised kernel: one CTA per SM walks 128-pixel tiles; 8 producer warps keep a ring of
// (kt, slot) stages full (they run ahead into the next tile), one thread streams the three weight tap blocks of each
// stage with cp.async.bulk, one thread issues the MMAs, 4 epilogue warps drain a double-buffered TMEM accumulator.
constexpr int CTP_STAGES = 2;
constexpr int CTP_PROD_WARPS = 8, CTP_EPI_WARPS = 4;
constexpr int CTP_THREADS = (CTP_PROD_WARPS + CTP_EPI_WARPS + 2) * 32;       // 448
constexpr int CTP_SMEM = CTP_STAGES * CT_STAGE + BM * 64 * 4 + 1024;          // ring + fp32 epilogue staging

__global__ void __launch_bounds__(CTP_THREADS, 1) conv_tap_tc_kernel(const GemmArgs g, const uint8_t* __restrict__ w_tc, int Mp /* B*T*(F+1) */) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[CTP_STAGES], empty_bar[CTP_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS/LDS, not generic ST/LD)
  const uint32_t smem_base = ptx::smem_u32(smem);
  uint8_t* stg_base = smem + CTP_STAGES * CT_STAGE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = g.Fin, F1 = F + 1;
  const int nst = g.taps_t * g.nslots;            // stages per tile: (kt, slot)
  const int ntiles = (Mp + BM - 1) / BM;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < CTP_STAGES; ++s) { ptx::mbar_init(&full_bar[s], CTP_PROD_WARPS + 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], CTP_EPI_WARPS * 32); }
    ptx::fence_barrier_init();
  }
  if (warp == CTP_PROD_WARPS + CTP_EPI_WARPS) ptx::tmem_alloc(&tmem_base_s, 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;         // accumulator ab at column 64 * ab

  if (warp < CTP_PROD_WARPS) {
    // ================= producers =================
    const int c = lane & 7, plane = (lane >> 3) & 1, rsub = warp * 2 + (lane >> 4);
    const uint32_t dst_lane = (uint32_t)(3 * CT_WTAP + plane * CT_APLANE + c * CT_LBO + rsub * 16);
    const int src_lane_off = plane * 128 + c * 16;
    constexpr int LAG = CTP_STAGES - 1;
    int gs = 0;                                   // global stage counter (runs across tiles)
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BM;
      long long pix[9];       // flat real pixel index of staged row j (centre time tap), -1 = zero row
      int trow[9];
#pragma unroll
      for (int p = 0; p < 9; ++p) {
        const int j = p * 16 + rsub;
        const int pp = m0 - 1 + j;
        pix[p] = -1; trow[p] = 0;
        if (j < CT_ROWS && pp >= 0 && pp < Mp) {
          const int bt = pp / F1, f = pp - bt * F1;
          if (f < F) { pix[p] = (long long)bt * F + f; trow[p] = bt % g.T; }
        }
      }
      for (int st = 0; st < nst; ++st, ++gs) {
        const int s = gs % CTP_STAGES;
        const uint32_t ph = (uint32_t)(gs / CTP_STAGES) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        const int kt = st / g.nslots, slot = st - kt * g.nslots;
        const int dt = (g.taps_t - 1 - kt) * g.dil;
        const uint8_t* src_base = reinterpret_cast<const uint8_t*>(g.a[slot]) + src_lane_off;
        const long long dpix = -(long long)dt * F;
        const uint32_t dst = smem_base + s * CT_STAGE + dst_lane;
#pragma unroll
        for (int p = 0; p < 9; ++p) {
          if (p * 16 + rsub < CT_ROWS) {
            const bool ok = pix[p] >= 0 && trow[p] >= dt;
            const long long q = ok ? pix[p] + dpix : 0;
            cp_async16_zfill(dst + p * 256, src_base + q * 256, ok ? 16u : 0u);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (gs >= LAG) {      // stage gs - LAG has landed: publish it
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&full_bar[(gs - LAG) % CTP_STAGES]);
        }
      }
    }
    if (gs > 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_bar[(gs - 1) % CTP_STAGES]);
    }
  } else if (warp < CTP_PROD_WARPS + CTP_EPI_WARPS) {
    // ================= epilogue warps: TMEM -> smem transpose -> coalesced bias store, padded index -> real pixel =================
    const int wq = warp & 3;                      // TMEM lane quarter (hardware: warp % 4); CTP_PROD_WARPS % 4 == 0
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
        const int pp = m0 + wq * 32 + R;
        if (pp < Mp) {
          const int bt = pp / F1, f = pp - bt * F1;
          if (f < F) {
            float4 val = stg[R * 16 + (cq ^ (R & 7))];
            val.x += bias.x; val.y += bias.y; val.z += bias.z; val.w += bias.w;
            st4(g.out + ((long long)bt * F + f) * g.ldo + cq * 4, val);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == CTP_PROD_WARPS + CTP_EPI_WARPS) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int gs = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int ab = it & 1;
        ptx::mbar_wait(&acc_empty[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 64);
        for (int st = 0; st < nst; ++st, ++gs) {
          const int s = gs % CTP_STAGES;
          const uint32_t ph = (uint32_t)(gs / CTP_STAGES) & 1u;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t base = smem_base + s * CT_STAGE;
          const uint32_t a_hi = base + 3 * CT_WTAP, a_lo = a_hi + CT_APLANE;
#pragma unroll
          for (int kf = 0; kf < 3; ++kf) {
            const uint64_t w_hi = ptx::umma_desc_sw128(base + kf * CT_WTAP);
            const uint64_t w_lo = ptx::umma_desc_sw128(base + kf * CT_WTAP + 64 * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t aoff = (uint32_t)(2 * k * CT_LBO + kf * 16);            // K-chunk pair 2k, 2k+1; pixel shift kf
              const uint64_t d_hi = umma_desc_noswz(a_hi + aoff, CT_LBO, 128);
              const uint64_t d_lo = umma_desc_noswz(a_lo + aoff, CT_LBO, 128);
              const uint64_t ko = (uint64_t)((k * 32) >> 4);
              ptx::mma_bf16(d_tmem, d_lo, w_hi + ko, IDESC, (st | kf | k) ? 1u : 0u);
              ptx::mma_bf16(d_tmem, d_hi, w_lo + ko, IDESC, 1u);
              ptx::mma_bf16(d_tmem, d_hi, w_hi + ko, IDESC, 1u);
            }
          }
          ptx::tc_commit(&empty_bar[s]);
        }
        ptx::tc_commit(&acc_full[ab]);
      }
    }
  } else {
    // ================= weight stager: the three taps of stage (kt, slot) =================
    if (lane == 0) {
      int gs = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int st = 0; st < nst; ++st, ++gs) {
          const int s = gs % CTP_STAGES;
          const uint32_t ph = (uint32_t)(gs / CTP_STAGES) & 1u;
          const int kt = st / g.nslots, slot = st - kt * g.nslots;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full_bar[s], 3 * CT_WTAP);
#pragma unroll
          for (int kf = 0; kf < 3; ++kf) {
            const int kc = (kt * 3 + kf) * g.nslots + slot;        // K order of the packed image: (kt, kf, slot)
            ptx::bulk_g2s(smem_base + s * CT_STAGE + kf * CT_WTAP, w_tc + (size_t)kc * CT_WTAP, CT_WTAP, &full_bar[s]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == CTP_PROD_WARPS + CTP_EPI_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

int launch_conv_tap(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  static int num_sms = 0;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(conv_tap_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CTP_SMEM);
    if (e != cudaSuccess) { set_error("conv tap: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    attr_done.set();
  }
  const long long Mp = (long long)s->B * s->T * (s->Fin + 1);
  SEB_REQUIRE(Mp < 2147483647LL - 256, SEB_EINVAL, "conv tap: padded pixel count overflows int");
  const long long ntiles = (Mp + BM - 1) / BM;
  dim3 grid((unsigned)(ntiles < num_sms ? ntiles : num_sms));
  conv_tap_tc_kernel<<<grid, CTP_THREADS, CTP_SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc), (int)Mp);
  SEB_CHECK_LAUNCH("conv_tap_tc_kernel");
  return 0;
}

bool conv_tap_enabled() {
  static int v = -1;
  // off by default: measured slower than conv_persist; SEB200_CONV_TAP=1 selects it
  if (v < 0) { const char* e = getenv("SEB200_CONV_TAP"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

}  // namespace seb
