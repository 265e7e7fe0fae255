// attention.cu -- multi-head self-attention core with Shaw relative positions (conformer.py:103-122), flash style:
// the (S, h, n, n) score / rel-pos / probability tensors of the reference are never materialised.
//
//   dots[i, j] = scale * q_i . (k_j + E[clamp(i - j, -512, 512) + 512]),   out_i = softmax_j(dots) . v
//
// variant 0: tensor-core kernel (mma.sync m16n8k8 TF32, fp32 accumulate, online softmax).  The rel-pos term of a
//            16 x 64 score tile is a second small GEMM  R = Q . E_window^T  (80 distinct offsets) followed by a
//            skewed read  S[i, j] += R[i, i - j - dlo]  through a warp-private shared-memory staging buffer.
//            TF32 (10-bit mantissa) is sufficient for the three attention contractions.
// variant 1: one-thread-per-query fp32 kernel; slow, used by the tests to cross-check variant 0.
#include "common.cuh"

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
// variant 0: TF32 mma.sync flash attention with the rel-pos GEMM + skew
// ------------------------------------------------------------------------------------------------
constexpr int AT_BQ = 64, AT_BK = 64, AT_LD = 20, AT_RLD = 84, AT_RW = 80;

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) attention_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ E,
                                                           const SebSeq sq, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int seq = blockIdx.x >> 2, h = blockIdx.x & 3, i0 = blockIdx.y * AT_BQ;
  const int n = sq.n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long base = seq_base(sq, seq);
  int e_lo = i0 - (n - 1); if (e_lo < -AT_MAXPOS) e_lo = -AT_MAXPOS;
  int e_hi = i0 + AT_BQ - 1; if (e_hi > AT_MAXPOS) e_hi = AT_MAXPOS;
  const int erows = e_hi - e_lo + 1;
  uint32_t* Es = reinterpret_cast<uint32_t*>(sm);          // [erows][20] tf32
  uint32_t* Ks = Es + erows * AT_LD;                        // [64][20]
  uint32_t* Vs = Ks + AT_BK * AT_LD;                        // [64][20]
  float* Rs = reinterpret_cast<float*>(Vs + AT_BK * AT_LD) + warp * 16 * AT_RLD;   // [16][84] per warp

  for (int idx = tid; idx < erows * 4; idx += 128) {
    const int r = idx >> 2, part = idx & 3;
    const float4 v = ldg4(E + (long long)(e_lo + AT_MAXPOS + r) * AT_D + part * 4);
    *reinterpret_cast<uint4*>(Es + r * AT_LD + part * 4) = make_uint4(f2tf32(v.x), f2tf32(v.y), f2tf32(v.z), f2tf32(v.w));
  }

  // Q fragments (pre-scaled by dim_head^-0.5 * log2(e): softmax evaluated with exp2)
  const int iw = i0 + warp * 16;
  const float qs = 0.25f * 1.4426950408889634f;
  uint32_t qa[2][4];
  {
    int r0 = iw + g, r1 = iw + g + 8;
    if (r0 >= n) r0 = n - 1;
    if (r1 >= n) r1 = n - 1;
    const float* q0 = qkv + (base + (long long)r0 * sq.pos_stride) * AT_ROW + h * AT_D;
    const float* q1 = qkv + (base + (long long)r1 * sq.pos_stride) * AT_ROW + h * AT_D;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      qa[ks][0] = f2tf32(__ldg(q0 + ks * 8 + t) * qs);
      qa[ks][1] = f2tf32(__ldg(q1 + ks * 8 + t) * qs);
      qa[ks][2] = f2tf32(__ldg(q0 + ks * 8 + t + 4) * qs);
      qa[ks][3] = f2tf32(__ldg(q1 + ks * 8 + t + 4) * qs);
    }
  }
  float o[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  float mrow[2] = {-1e30f, -1e30f}, lrow[2] = {0.f, 0.f};

  for (int j0 = 0; j0 < n; j0 += AT_BK) {
    __syncthreads();
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int idx = tid + rep * 128, key = idx >> 2, part = idx & 3, j = j0 + key;
      float4 kv = make_float4(0, 0, 0, 0), vv = kv;
      if (j < n) {
        const float* p = qkv + (base + (long long)j * sq.pos_stride) * AT_ROW + h * AT_D + part * 4;
        kv = ldg4(p + 64); vv = ldg4(p + 128);
      }
      *reinterpret_cast<uint4*>(Ks + key * AT_LD + part * 4) = make_uint4(f2tf32(kv.x), f2tf32(kv.y), f2tf32(kv.z), f2tf32(kv.w));
      *reinterpret_cast<uint4*>(Vs + key * AT_LD + part * 4) = make_uint4(f2tf32(vv.x), f2tf32(vv.y), f2tf32(vv.z), f2tf32(vv.w));
    }
    __syncthreads();

    // content scores S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t* kr = Ks + (nt * 8 + g) * AT_LD + ks * 8 + t;
        mma_tf32(s[nt], qa[ks], kr[0], kr[4]);
      }
    }
    // relative-position scores R[r, dd] = q_r . E[clamp(dlo + dd)],  dd in [0, 80)
    const int dlo = iw - j0 - 64;
#pragma unroll
    for (int nt = 0; nt < AT_RW / 8; ++nt) {
      int d = dlo + nt * 8 + g;
      d = d < e_lo ? e_lo : (d > e_hi ? e_hi : d);
      const uint32_t* er = Es + (d - e_lo) * AT_LD + t;
      float r4[4] = {0.f, 0.f, 0.f, 0.f};
      mma_tf32(r4, qa[0], er[0], er[4]);
      mma_tf32(r4, qa[1], er[8], er[12]);
      *reinterpret_cast<float2*>(Rs + g * AT_RLD + nt * 8 + 2 * t) = make_float2(r4[0], r4[1]);
      *reinterpret_cast<float2*>(Rs + (g + 8) * AT_RLD + nt * 8 + 2 * t) = make_float2(r4[2], r4[3]);
    }
    __syncwarp();
    // skew: S[r, c] += R[r, 64 + r - c]; mask keys beyond the sequence
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = g + ((e >> 1) << 3), c = nt * 8 + 2 * t + (e & 1);
        const float v = s[nt][e] + Rs[r * AT_RLD + 64 + r - c];
        s[nt][e] = (j0 + c < n) ? v : -1e30f;
      }
    }
    __syncwarp();
    // online softmax (rows g and g+8 of this warp's slab)
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      float mx = -1e30f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * rh], s[nt][2 * rh + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float mn = fmaxf(mrow[rh], mx);
      const float corr = exp2f(mrow[rh] - mn);
      mrow[rh] = mn;
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = exp2f(s[nt][2 * rh] - mn), p1 = exp2f(s[nt][2 * rh + 1] - mn);
        s[nt][2 * rh] = p0; s[nt][2 * rh + 1] = p1;
        sum += p0 + p1;
      }
      lrow[rh] = lrow[rh] * corr + sum;
      o[0][2 * rh] *= corr; o[0][2 * rh + 1] *= corr;
      o[1][2 * rh] *= corr; o[1][2 * rh + 1] *= corr;
    }
    // O += P V.  The C-fragment columns (2t, 2t+1) of an 8-key block are fed as A-fragment k-slots (t, t+4); V rows
    // are read with the same key permutation, so the sum over keys is unchanged.
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      uint32_t pa[4] = {f2tf32(s[kb][0]), f2tf32(s[kb][2]), f2tf32(s[kb][1]), f2tf32(s[kb][3])};
      const uint32_t* v0 = Vs + (kb * 8 + 2 * t) * AT_LD + g;
      mma_tf32(o[0], pa, v0[0], v0[AT_LD]);
      mma_tf32(o[1], pa, v0[8], v0[AT_LD + 8]);
    }
  }
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    float l = lrow[rh];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int i = iw + g + 8 * rh;
    if (i < n) {
      const float inv = 1.0f / l;
      float* op = out + (base + (long long)i * sq.pos_stride) * 64 + h * AT_D + 2 * t;
      *reinterpret_cast<float2*>(op) = make_float2(o[0][2 * rh] * inv, o[0][2 * rh + 1] * inv);
      *reinterpret_cast<float2*>(op + 8) = make_float2(o[1][2 * rh] * inv, o[1][2 * rh + 1] * inv);
    }
  }
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_attention(const float* qkv, const float* rel_pos_emb, const SebSeq* seq, float* out, int variant, void* stream) {
  SEB_REQUIRE(qkv && rel_pos_emb && seq && out && aligned16(qkv) && aligned16(out) && aligned16(rel_pos_emb), SEB_EINVAL, "attention: null/unaligned argument");
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
    attention_simt_kernel<<<grid, ATS_BQ, smem, st>>>(qkv, rel_pos_emb, *seq, out);
    SEB_CHECK_LAUNCH("attention_simt_kernel");
    return 0;
  }
  int erows = n + AT_BQ - 1; if (erows > 2 * AT_MAXPOS + 1) erows = 2 * AT_MAXPOS + 1;
  const int smem = (erows * AT_LD + 2 * AT_BK * AT_LD + 4 * 16 * AT_RLD) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid(seq->nseq * AT_H, (n + AT_BQ - 1) / AT_BQ);
  SEB_REQUIRE(grid.y <= 65535u, SEB_EINVAL, "attention: sequence too long");
  attention_mma_kernel<<<grid, 128, smem, st>>>(qkv, rel_pos_emb, *seq, out);
  SEB_CHECK_LAUNCH("attention_mma_kernel");
  return 0;
}
