// dwpw2.cu -- second half of the conformer convolution module in ONE kernel (conformer.py:166-169 + the block residual :204):
//
//     v   = Swish( BatchNorm1d_eval( DepthWiseConv1d_k31( u ) ) )            along the sequence axis, 128 channels
//     out = resid + W3 . v + b3                                              pointwise Conv1d 128 -> 64
//
// The standalone pair (dwconv_bn_swish_kernel, then the 128 -> 64 token GEMM) wrote v [tokens, 128] fp32 to HBM and read
// it back: 4.2 GB per conformer block for nothing.  Here the depthwise result never leaves the SM: the 256 depthwise threads
// write it as bf16 hi|lo straight into the UMMA K-major SWIZZLE_128B operand tile in shared memory, one thread issues the
// six tcgen05.mma K-steps (3-product split) against the resident W3 image, and four epilogue warps add bias + residual and
// store coalesced.
//
// Work item = 128 consecutive positions of one sequence, processed as two 64-position halves so that the cp.async staging
// of the next half ((64 + 30) x 128 fp32 = 47 KB, issued by the 512 depthwise threads themselves: six 16-byte copies each;
// 94 separate 512-byte cp.async.bulk per half from one thread cost ~46 cycles apiece in the TMA unit and starved the
// arithmetic) overlaps the depthwise arithmetic of the current one.  Persistent:
// one CTA per SM walks the (sequence, chunk) list; chunks are the fast index, so the 30 halo rows shared by neighbouring
// chunks are served by L2.
//   warps 0-15  depthwise (thread = one channel x 16 positions, the 31 taps and a 23-row window in registers: ~80 registers, so
//               16 such warps fit -- 8 warps of channel PAIRS (124 registers) left the FMA pipe latency-bound at 2 warps per scheduler)
//   warps 16-23 epilogue  (lane quarter x column half: residual rows prefetched, TMEM -> warp-private smem transpose ->
//               + b3 + resid -> coalesced store; with 4 warps the two serial column passes per item paced the kernel)
//   warp 24     loads W3 once, then issues the MMAs
// STATUS: correct (tests/test_gpu_kernels.py::test_dwconv_pw2_fused) but NOT the default path: at configs[1] it takes 2.6 ms per
// launch against 1.05 + 0.80 ms for the two standalone kernels.  The standalone depthwise kernel runs 16 warps per SM at 118
// registers (31 taps + window as packed pairs in registers); next to epilogue / MMA warps in one CTA the register file only
// leaves 72-80 per thread, and every organisation tried (8 warps of channel pairs, 16 warps with shared-memory taps, 16 warps
// of single channels) ends issue- or latency-bound.  A setmaxnreg split of the register file is the next thing to try.
#include "gemm_engine.cuh"

namespace seb {

constexpr int DP_TI = 64, DP_K = 31, DP_PAD = 15, DP_C = 128;
constexpr int DP_ROWS = DP_TI + DP_K - 1;                  // 94 staged rows per half
constexpr int DP_USTAGE = DP_ROWS * DP_C * 4;              // 48,128 B
constexpr int DP_DW_WARPS = 16, DP_EPI_WARPS = 8;
constexpr int DP_W_EPI0 = DP_DW_WARPS;                     // 16 (multiple of 4: warp % 4 = TMEM lane quarter)
constexpr int DP_W_MMA = DP_W_EPI0 + DP_EPI_WARPS;         // 24 (also loads W3 once)
constexpr int DP_THREADS = (DP_W_MMA + 1) * 32;            // 800: 80 registers per thread
constexpr int DP_VPLANE = BM * 128;                        // 16 KB: one bf16 plane of a 128 x 64 operand chunk
constexpr int DP_SMEM = 1024 + 2 * DP_USTAGE + 4 * DP_VPLANE + 2 * 2 * 64 * 128 + DP_EPI_WARPS * 4096;

struct DwPw2Args {
  const float* u; SebSeq sq;
  const float* w; const float* bn_scale; const float* bn_shift;
  const uint8_t* w3; const float* b3;
  const float* resid; float* out;
  int nchunks; long long nitems;
};

__device__ __forceinline__ float2 dp_ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__global__ void __launch_bounds__(DP_THREADS, 1) dwpw2_kernel(const DwPw2Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t v_full, v_empty, acc_full[2], acc_empty[2], w_full;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sV = smem;                                   // [kc 2][hi | lo][128 rows x 128 B]  64 KB (1024-aligned)
  uint8_t* sW = sV + 4 * DP_VPLANE;                     // [kc 2][hi | lo][64 rows x 128 B]   32 KB
  uint8_t* sU = sW + 2 * 2 * 64 * 128;                  // [2][94 rows][128 fp32]             94 KB
  uint8_t* sStg = sU + 2 * DP_USTAGE;                   // 4 KB per epilogue warp
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.sq.n;
  const long long my_items = ((long long)blockIdx.x < a.nitems) ? (a.nitems - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], DP_EPI_WARPS * 32);
    }
    ptx::mbar_init(&v_full, DP_DW_WARPS * 32); ptx::mbar_init(&v_empty, 1); ptx::mbar_init(&w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == DP_W_MMA) ptx::tmem_alloc(&tmem_base_s, 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;               // accumulator ab at column 64 * ab

  if (warp < DP_DW_WARPS) {
    // ================= depthwise + BN + Swish -> bf16 hi|lo rows of the GEMM operand =================
    const int c = tid & 127, ph = tid >> 7;             // channel, 16-position quarter of the half
    float wr[DP_K];
#pragma unroll
    for (int k = 0; k < DP_K; ++k) wr[k] = __ldg(a.w + k * DP_C + c);
    const float sc = __ldg(a.bn_scale + c), sh = __ldg(a.bn_shift + c);
    // operand address of this channel inside a 128-byte row: k-chunk c / 64, 16-byte chunk (c % 64) / 8, byte (c % 8) * 2
    const int vch = (c & 63) >> 3;
    uint8_t* vbase = sV + (c >> 6) * 2 * DP_VPLANE + (c & 7) * 2;
    // staging of half g (g = 2 * local item + half) into slot g & 1: 94 rows x 32 chunks of 16 B over 512 threads
    const long long row_bytes = a.sq.pos_stride * DP_C * 4;
    const long long nhalves = 2 * my_items;
    auto stage = [&](long long g) {
      const long long item = blockIdx.x + (g >> 1) * gridDim.x;
      const int seq = (int)(item / a.nchunks), chunk = (int)(item % a.nchunks);
      const int h0 = chunk * BM + (int)(g & 1) * DP_TI;                // first output position of the half
      if (h0 < n) {
        const long long base = (long long)(seq / a.sq.inner) * a.sq.outer_stride + (seq % a.sq.inner);
        const uint8_t* seq0 = reinterpret_cast<const uint8_t*>(a.u + base * DP_C) + lane * 16;
        const uint32_t dst = ptx::smem_u32(sU) + (uint32_t)(g & 1) * DP_USTAGE + lane * 16;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int r = warp + 16 * k;
          if (r < DP_ROWS) {
            const int i = h0 - DP_PAD + r;
            const bool ok = i >= 0 && i < n;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + r * (DP_C * 4)), "l"(ok ? seq0 + (long long)i * row_bytes : seq0), "r"(ok ? 16u : 0u) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nhalves > 0) stage(0);
    for (long long it = 0; it < my_items; ++it) {
      const long long item = blockIdx.x + it * gridDim.x;
      const int chunk = (int)(item % a.nchunks);
      const int i0 = chunk * BM;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const long long g = 2 * it + half;
        const int s = (int)(g & 1);
        // everyone has finished reading slot s ^ 1 (half g - 1): refill it with half g + 1, then wait for half g
        asm volatile("bar.sync 1, %0;" ::"n"(DP_DW_WARPS * 32) : "memory");
        if (g + 1 < nhalves) { stage(g + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(DP_DW_WARPS * 32) : "memory");
        const float* t1 = reinterpret_cast<const float*>(sU + s * DP_USTAGE) + c;          // row stride DP_C floats
        const bool live = i0 + half * DP_TI < n;
        uint32_t vh[8], vl[8];                           // 16 outputs as packed (position o, o + 1) bf16 pairs: hi and lo planes
        if (live) {
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            const int gpos = ph * 16 + gi * 8;          // first output position of this group inside the half
            float acc[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] = 0.f;
            {   // taps 0..15 use window rows gpos .. gpos+22
              float win[23];
#pragma unroll
              for (int r = 0; r < 23; ++r) win[r] = t1[(gpos + r) * DP_C];
#pragma unroll
              for (int k = 0; k < 16; ++k)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(wr[k], win[o + k], acc[o]);
            }
            {   // taps 16..30 use window rows gpos+16 .. gpos+37
              float win[22];
#pragma unroll
              for (int r = 0; r < 22; ++r) win[r] = t1[(gpos + 16 + r) * DP_C];
#pragma unroll
              for (int k = 16; k < DP_K; ++k)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(wr[k], win[o + k - 16], acc[o]);
            }
#pragma unroll
            for (int o = 0; o < 8; o += 2) {
              float v0 = fmaf(acc[o], sc, sh), v1 = fmaf(acc[o + 1], sc, sh);
              v0 *= sigmoidf_acc(v0); v1 *= sigmoidf_acc(v1);
              split_bf16x2(v0, v1, vh[gi * 4 + (o >> 1)], vl[gi * 4 + (o >> 1)]);
            }
          }
        }
        if (half == 0) ptx::mbar_wait(&v_empty, ((uint32_t)it & 1u) ^ 1u);   // the MMAs of the previous item have consumed sV
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = half * DP_TI + ph * 16 + 2 * j;            // operand row = position inside the 128-position chunk
            uint8_t* p0 = vbase + row * 128 + ((vch ^ (row & 7)) << 4);
            uint8_t* p1 = vbase + (row + 1) * 128 + ((vch ^ ((row + 1) & 7)) << 4);
            *reinterpret_cast<uint16_t*>(p0) = (uint16_t)(vh[j] & 0xffffu);
            *reinterpret_cast<uint16_t*>(p1) = (uint16_t)(vh[j] >> 16);
            *reinterpret_cast<uint16_t*>(p0 + DP_VPLANE) = (uint16_t)(vl[j] & 0xffffu);
            *reinterpret_cast<uint16_t*>(p1 + DP_VPLANE) = (uint16_t)(vl[j] >> 16);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&v_full);
    }
  } else if (warp < DP_W_MMA) {
    // ================= epilogue: out = resid + acc + b3, coalesced through a warp-private transpose =================
    const int wq = warp & 3, c0 = ((warp - DP_W_EPI0) >> 2) * 32;   // TMEM lane quarter (rows 32 wq ..), column half
    float4* stg = reinterpret_cast<float4*>(sStg + (warp - DP_W_EPI0) * 4096);      // [32 rows][8 x float4]
    const int ch = lane & 7;
    const float4 bb = ldg4(a.b3 + c0 + ch * 4);
    for (long long it = 0; it < my_items; ++it) {
      const long long item = blockIdx.x + it * gridDim.x;
      const int seq = (int)(item / a.nchunks), chunk = (int)(item % a.nchunks);
      const int i0 = chunk * BM;
      const long long base = (long long)(seq / a.sq.inner) * a.sq.outer_stride + (seq % a.sq.inner);
      const int ab = (int)(it & 1);
      float4 res[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {               // residual rows before the accumulator wait: `out` aliases `resid`
        const int i = i0 + wq * 32 + i8 * 4 + (lane >> 3);
        res[i8] = (i < n) ? *reinterpret_cast<const float4*>(a.resid + (base + (long long)i * a.sq.pos_stride) * 64 + c0 + ch * 4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ptx::mbar_wait(&acc_full[ab], (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * 64 + c0);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float v[8];
        ptx::tmem_ld8(taddr + j, v);
        stg[lane * 8 + (((j >> 2) + 0) ^ (lane & 7))] = make_float4(v[0], v[1], v[2], v[3]);
        stg[lane * 8 + (((j >> 2) + 1) ^ (lane & 7))] = make_float4(v[4], v[5], v[6], v[7]);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acc_empty[ab]);
      __syncwarp();
      float4 vals[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int R = i8 * 4 + (lane >> 3);
        vals[i8] = stg[R * 8 + (ch ^ (R & 7))];
      }
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int i = i0 + wq * 32 + i8 * 4 + (lane >> 3);
        if (i < n) {
          float4 v = vals[i8];
          v.x += bb.x + res[i8].x; v.y += bb.y + res[i8].y; v.z += bb.z + res[i8].z; v.w += bb.w + res[i8].w;
          st4(a.out + (base + (long long)i * a.sq.pos_stride) * 64 + c0 + ch * 4, v);
        }
      }
      __syncwarp();
    }
  } else if (warp == DP_W_MMA) {
    // ================= MMA issuer: ACC[ab] = V . W3^T (K = 128 in two 64-wide chunks, 3-product split) =================
    if (lane == 0 && my_items > 0) {
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t uV = ptx::smem_u32(sV), uW = ptx::smem_u32(sW);
      ptx::mbar_arrive_expect_tx(&w_full, 2 * 2 * 64 * 128);
      ptx::bulk_g2s(uW, a.w3, 16384, &w_full);
      ptx::bulk_g2s(uW + 16384, a.w3 + 16384, 16384, &w_full);
      ptx::mbar_wait(&w_full, 0);
      for (long long it = 0; it < my_items; ++it) {
        const int ab = (int)(it & 1);
        ptx::mbar_wait(&acc_empty[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        ptx::mbar_wait(&v_full, (uint32_t)it & 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 64);
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint64_t a_hi = ptx::umma_desc_sw128(uV + kc * 2 * DP_VPLANE), a_lo = ptx::umma_desc_sw128(uV + kc * 2 * DP_VPLANE + DP_VPLANE);
          const uint64_t w_hi = ptx::umma_desc_sw128(uW + kc * 2 * 64 * 128), w_lo = ptx::umma_desc_sw128(uW + kc * 2 * 64 * 128 + 64 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            ptx::mma_bf16(d_tmem, a_lo + ko, w_hi + ko, IDESC, (kc | k) ? 1u : 0u);
            ptx::mma_bf16(d_tmem, a_hi + ko, w_lo + ko, IDESC, 1u);
            ptx::mma_bf16(d_tmem, a_hi + ko, w_hi + ko, IDESC, 1u);
          }
        }
        ptx::tc_commit(&v_empty);
        ptx::tc_commit(&acc_full[ab]);
      }
    }
  }
  __syncthreads();
  if (warp == DP_W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_dwconv_pw2(const float* u, const SebSeq* seq, const float* w, const float* bn_scale, const float* bn_shift,
                                 const void* w3_tc, const float* b3, const float* resid, float* out, void* stream) {
  SEB_REQUIRE(u && seq && w && bn_scale && bn_shift && w3_tc && b3 && resid && out, SEB_EINVAL, "dwconv_pw2: null argument");
  SEB_REQUIRE(aligned16(u) && aligned16(w3_tc) && aligned16(b3) && aligned16(resid) && aligned16(out), SEB_EALIGN, "dwconv_pw2: unaligned pointer");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0, SEB_EINVAL, "dwconv_pw2: bad sequence descriptor");
  static PerDeviceOnce attr_done;
  static int num_sms = 0;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(dwpw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM);
    if (e != cudaSuccess) { set_error("dwconv_pw2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    attr_done.set();
  }
  DwPw2Args a;
  a.u = u; a.sq = *seq; a.w = w; a.bn_scale = bn_scale; a.bn_shift = bn_shift;
  a.w3 = reinterpret_cast<const uint8_t*>(w3_tc); a.b3 = b3; a.resid = resid; a.out = out;
  a.nchunks = (seq->n + BM - 1) / BM;
  a.nitems = (long long)seq->nseq * a.nchunks;
  const unsigned grid = (unsigned)(a.nitems < num_sms ? a.nitems : num_sms);
  dwpw2_kernel<<<grid, DP_THREADS, DP_SMEM, (cudaStream_t)stream>>>(a);
  SEB_CHECK_LAUNCH("dwpw2_kernel");
  return 0;
}
