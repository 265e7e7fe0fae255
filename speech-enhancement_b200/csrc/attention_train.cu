// attention_train.cu -- the attention core with Shaw relative positions (conformer.py:103-122) for the TRAINING step:
// an fp32-grade forward that also returns the log-sum-exp of every (token, head) row, and its backward
//
//   S = scale * (Q K^T + skew(Q E_win^T)),  P = softmax(S),  O = P V
//   dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(dO o O));  dQ = scale * (dS K + unskew(dS) E_win);  dK = scale * dS^T Q;
//   dE[r] = scale * sum over (sequence, head, i, j : clamp(i - j) = r) of dS[i, j] q_i            (the 1025 x 16 table under the clamp)
//
// flash style (no (S, h, n, n) tensor), 64 x 64 tiles, 4 warps per CTA, every contraction on mma.sync m16n8k8 with the 3xTF32 split of
// mma_tf32.cuh (fp32 range, ~2^-21 operand error).  Two backward kernels, no atomics on global memory, deterministic:
//   * query-major  (CTA = 64 queries, loops over key tiles):  dQ and dE.  The rel-pos gradient of a tile is the banded matrix
//     dR[i, w] = dS[i, i - j + 63] scattered into shared memory; dQ += dR E_win and dE_win = dR^T Q are two more small GEMMs; dE_win is
//     added into a CTA-resident copy of the table rows a length-n sequence can address (41 KB of shared memory at n = 321, 64 KB from
//     n = 513), flushed once per persistent CTA and summed by a finish kernel.
//   * key-major    (CTA = 64 keys, loops over query tiles):   dK and dV from the transposed tile S^T = K Q^T (+ the same R tile).
// qkv is the fp32 projection [tokens, 192] = (q | k | v), heads 4 x 16; token addressing through SebSeq like every conformer kernel.
#include "mma_tf32.cuh"
#include <type_traits>

namespace seb {

constexpr int TA_B = 64, TA_D = 16, TA_LD = 20, TA_RLD = 132, TA_PLD = 68, TA_MAXPOS = 512, TA_ROW = 192, TA_EROWS = 2 * TA_MAXPOS + 1;
constexpr float TA_SCALE = 0.25f, TA_NEG = -1e30f;

__device__ __forceinline__ long long ta_seq_base(const SebSeq& sq, int seq) {
  return (long long)(seq / sq.inner) * sq.outer_stride + (seq % sq.inner);
}
// rows r0 .. r0 + 63 of a 16-wide column block (column offset `col`) of a [tokens, ld] matrix -> dst[64][TA_LD]; rows >= n are zero
__device__ __forceinline__ void ta_load_tile(float* dst, const float* src, long long base, long long pos_stride, int ld, int col, int r0, int n) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = threadIdx.x + u * 128, r = idx >> 2, part = idx & 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n) v = ldg4(src + (base + (long long)(r0 + r) * pos_stride) * ld + col + part * 4);
    *reinterpret_cast<float4*>(dst + r * TA_LD + part * 4) = v;
  }
}
// E window of the tile (i0, j0): row w = E[clamp(i0 - j0 - 63 + w)], w < 127; row 127 = 0
__device__ __forceinline__ void ta_load_window(float* Es, const float* E, int dbase) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int idx = threadIdx.x + u * 128, w = idx >> 2, part = idx & 3;
    int d = dbase + w;
    d = d < -TA_MAXPOS ? -TA_MAXPOS : (d > TA_MAXPOS ? TA_MAXPOS : d);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w < 127) v = ldg4(E + (long long)(d + TA_MAXPOS) * TA_D + part * 4);
    *reinterpret_cast<float4*>(Es + w * TA_LD + part * 4) = v;
  }
}
// R rows of this warp: Rs[16 warp + r][w] = q_r . E_win[w] for the window columns its 16 queries can address: a query row il reads
// columns il - jl + 63 in [il, il + 63], so the warp needs w in [16 warp, 16 warp + 79] -- 10 of the 16 n-tiles
constexpr int TA_BAND = 80;
// FR: the key tile has only nl < 64 live keys (local j < nl), so a row reads columns il - jl + 63 >= il + 64 - nl: band n-tiles below nt_first = (64 - nl) / 8 are never read
template <bool FR>
__device__ __forceinline__ void ta_rel_rows(const float* Qs, const float* Es, float* Rs, int warp, int nt_first) {
  float acc[1][10][4];
  tf32::zero(acc);
  tf32::warp_gemm_fr<FR, 1, 10>(acc, 2, 0, 2, nt_first, 10, [&](int r, int k) { return Qs[(warp * 16 + r) * TA_LD + k]; },
                                [&](int k, int c) { return Es[(warp * 16 + c) * TA_LD + k]; });
#pragma unroll
  for (int nt = 0; nt < 10; ++nt)
    if (!FR || nt >= nt_first) {
#pragma unroll
      for (int e = 0; e < 4; ++e) Rs[(warp * 16 + tf32::c_row(0, e)) * TA_RLD + warp * 16 + tf32::c_col(nt, e)] = acc[0][nt][e];
    }
}
// live keys / rows of the 64-wide tile at t0 and the n-tile counts that go with them
__device__ __forceinline__ int ta_live(int n, int t0) { return n - t0 < TA_B ? n - t0 : TA_B; }

// ------------------------------------------------------------------------------------------------------------------------------------
// forward: out [tokens, 64], lse [tokens, 4]
// ------------------------------------------------------------------------------------------------------------------------------------
constexpr int TAF_SMEM = (3 * TA_B * TA_LD + 128 * TA_LD + TA_B * TA_RLD) * 4;

__global__ void __launch_bounds__(128) attention_train_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ E, const SebSeq sq, int nqt,
                                                                 float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) float ta_sm[];
  float* Qs = ta_sm; float* Ks = Qs + TA_B * TA_LD; float* Vs = Ks + TA_B * TA_LD; float* Es = Vs + TA_B * TA_LD;
  float* Rs = Es + 128 * TA_LD;
  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x % nqt, sh = blockIdx.x / nqt, h = sh & 3, seq = sh >> 2;
  const int n = sq.n, i0 = qt * TA_B;
  const long long base = ta_seq_base(sq, seq);
  ta_load_tile(Qs, qkv, base, sq.pos_stride, TA_ROW, h * TA_D, i0, n);
  float m[2] = {TA_NEG, TA_NEG}, l[2] = {0.f, 0.f};
  float o[1][2][4];
  tf32::zero(o);
  const bool warp_live = i0 + warp * 16 < n;             // a warp whose 16 query rows all lie past the sequence only keeps the barriers
  // one key tile; FR: fewer than 64 live keys (n = 64 k + 1, the model's frame counts, leaves ONE in the last tile) -- the dead n-tiles / k-steps are skipped
  auto tile = [&](auto fr_tag, int j0) {
    constexpr bool FR = decltype(fr_tag)::value;
    const int nl = ta_live(n, j0), nth = (nl + 7) >> 3, ntf = (TA_B - nl) >> 3;
    ta_rel_rows<FR>(Qs, Es, Rs, warp, ntf);
    __syncwarp();
    float s[1][8][4];
    tf32::zero(s);
    tf32::warp_gemm_fr<FR, 1, 8>(s, 2, 0, 2, 0, nth, [&](int r, int k) { return Qs[(warp * 16 + r) * TA_LD + k]; }, [&](int k, int c) { return Ks[c * TA_LD + k]; });
    float mx[2] = {TA_NEG, TA_NEG};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (!FR || nt < nth) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int il = warp * 16 + tf32::c_row(0, e), jl = tf32::c_col(nt, e);
          float v = TA_SCALE * (s[0][nt][e] + Rs[il * TA_RLD + il - jl + 63]);
          if (j0 + jl >= n) v = TA_NEG;
          s[0][nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m[r], mx[r]);
      corr[r] = expf(m[r] - mn);
      m[r] = mn;
      l[r] *= corr[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (!FR || nt < nth) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = expf(s[0][nt][e] - m[e >> 1]);
          l[e >> 1] += p;
          const int il = warp * 16 + tf32::c_row(0, e);
          Rs[il * TA_RLD + il - tf32::c_col(nt, e) + 63] = p;          // P[i, j] replaces the R entry it was built from (same skewed slot)
        }
      }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[0][nt][e] *= corr[e >> 1];
    __syncwarp();
    tf32::warp_gemm_fr<FR, 1, 2>(o, 8, 0, nth, 0, 2, [&](int r, int k) { return Rs[(warp * 16 + r) * TA_RLD + warp * 16 + r - k + 63]; },
                                 [&](int k, int c) { return Vs[k * TA_LD + c]; });
  };
  for (int j0 = 0; j0 < n; j0 += TA_B) {
    __syncthreads();
    ta_load_tile(Ks, qkv, base, sq.pos_stride, TA_ROW, 64 + h * TA_D, j0, n);
    ta_load_tile(Vs, qkv, base, sq.pos_stride, TA_ROW, 128 + h * TA_D, j0, n);
    ta_load_window(Es, E, i0 - j0 - 63);
    __syncthreads();
    if (warp_live) {
      if (n - j0 >= TA_B) tile(std::false_type{}, j0); else tile(std::true_type{}, j0);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; e += 2) {
      const int i = i0 + warp * 16 + tf32::c_row(0, e);
      if (i < n) {
        const float inv = 1.0f / l[e >> 1];
        const long long tok = base + (long long)i * sq.pos_stride;
        *reinterpret_cast<float2*>(out + tok * 64 + h * TA_D + tf32::c_col(nt, e)) = make_float2(o[0][nt][e] * inv, o[0][nt][e + 1] * inv);
        if (nt == 0 && (threadIdx.x & 3) == 0) lse[tok * 4 + h] = m[e >> 1] + logf(l[e >> 1]);
      }
    }
}

// D[token][h] = sum_d dO * O over the head's 16 channels
__global__ void __launch_bounds__(256) attention_rowdot_kernel(const float* __restrict__ o, const float* __restrict__ d_o, long long rows4, float* __restrict__ D) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows4; i += (long long)gridDim.x * 256) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 a = ldg4(o + i * 16 + c * 4), b = ldg4(d_o + i * 16 + c * 4);
      s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    D[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// backward, query-major: dQ (into dqkv[:, 0:64]) and the per-CTA partial of dE
// ------------------------------------------------------------------------------------------------------------------------------------
// shared memory: four 64 x 16 tiles, the E window, the R / dR tile, per-row lse / D, and the CTA's copy of the rel-pos gradient table --
// only the rows a length-n sequence can address: offsets clamp(i - j) in [-min(n - 1, 512), +min(n - 1, 512)]
__host__ __device__ constexpr int taq_rows(int n) { return 2 * (n - 1 < TA_MAXPOS ? n - 1 : TA_MAXPOS) + 1; }
__host__ __device__ constexpr int taq_smem(int n) { return (4 * TA_B * TA_LD + 128 * TA_LD + TA_B * TA_RLD + 2 * TA_B + taq_rows(n) * TA_D) * 4; }

__global__ void __launch_bounds__(128) attention_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ E, const SebSeq sq, int nqt, long long nitems,
                                                             const float* __restrict__ lse, const float* __restrict__ Dg, const float* __restrict__ d_o,
                                                             float* __restrict__ dqkv, float* __restrict__ de_partial) {
  extern __shared__ __align__(16) float ta_sm[];
  float* Qs = ta_sm; float* dOs = Qs + TA_B * TA_LD; float* Ks = dOs + TA_B * TA_LD; float* Vs = Ks + TA_B * TA_LD; float* Es = Vs + TA_B * TA_LD;
  float* Rs = Es + 128 * TA_LD; float* Ls = Rs + TA_B * TA_RLD; float* Ds = Ls + TA_B; float* dEt = Ds + TA_B;
  const int warp = threadIdx.x >> 5;
  const int n = sq.n;
  const int emax = n - 1 < TA_MAXPOS ? n - 1 : TA_MAXPOS, trows = 2 * emax + 1;       // table row of offset d: d + emax
  for (int i = threadIdx.x; i < trows * TA_D; i += 128) dEt[i] = 0.f;
  for (int i = threadIdx.x; i < TA_B * TA_RLD; i += 128) Rs[i] = 0.f;                  // entries outside a warp's band are never written again: they stay zero
  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int qt = (int)(item % nqt), sh = (int)(item / nqt), h = sh & 3, seq = sh >> 2;
    const int i0 = qt * TA_B;
    const long long base = ta_seq_base(sq, seq);
    __syncthreads();
    ta_load_tile(Qs, qkv, base, sq.pos_stride, TA_ROW, h * TA_D, i0, n);
    ta_load_tile(dOs, d_o, base, sq.pos_stride, 64, h * TA_D, i0, n);
    if (threadIdx.x < TA_B) {
      const int i = i0 + threadIdx.x;
      const long long tok = base + (long long)(i < n ? i : 0) * sq.pos_stride;
      Ls[threadIdx.x] = i < n ? lse[tok * 4 + h] : 0.f;
      Ds[threadIdx.x] = i < n ? Dg[tok * 4 + h] : 0.f;
    }
    float dq[1][2][4];
    tf32::zero(dq);
    const int nr = ta_live(n, i0);                      // live query rows of this tile
    const bool warp_live = warp * 16 < nr;              // a warp whose 16 query rows all lie past the sequence skips its row-owned work
    // the row-owned part of one key tile; FR: fewer than 64 live keys -- the dead n-tiles / k-steps are skipped
    auto tile_rows = [&](auto fr_tag, int j0) {
      constexpr bool FR = decltype(fr_tag)::value;
      const int nl = ta_live(n, j0), nth = (nl + 7) >> 3, ntf = (TA_B - nl) >> 3;
      ta_rel_rows<FR>(Qs, Es, Rs, warp, ntf);
      __syncwarp();
      float s[1][8][4], dp[1][8][4];
      tf32::zero(s);
      tf32::zero(dp);
      tf32::warp_gemm_fr<FR, 1, 8>(s, 2, 0, 2, 0, nth, [&](int r, int k) { return Qs[(warp * 16 + r) * TA_LD + k]; }, [&](int k, int c) { return Ks[c * TA_LD + k]; });
      tf32::warp_gemm_fr<FR, 1, 8>(dp, 2, 0, 2, 0, nth, [&](int r, int k) { return dOs[(warp * 16 + r) * TA_LD + k]; }, [&](int k, int c) { return Vs[c * TA_LD + k]; });
      // dS' = scale * P * (dP - D); keep it in s
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if (!FR || nt < nth) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int rl = warp * 16 + tf32::c_row(0, e), jl = tf32::c_col(nt, e);
            const float v = TA_SCALE * (s[0][nt][e] + Rs[rl * TA_RLD + rl - jl + 63]);
            const float p = (j0 + jl < n && i0 + rl < n) ? expf(v - Ls[rl]) : 0.f;
            s[0][nt][e] = TA_SCALE * p * (dp[0][nt][e] - Ds[rl]);
          }
        }
      __syncwarp();                                     // every lane has read its R entries: the rows can be overwritten by dR
      // own rows of dR (the banded matrix dR[i, w] = dS'[i, i - w + 63]): clear the warp's band, then scatter
      for (int idx = threadIdx.x & 31; idx < 16 * (TA_BAND / 4); idx += 32)
        *reinterpret_cast<float4*>(Rs + (warp * 16 + idx / (TA_BAND / 4)) * TA_RLD + warp * 16 + (idx % (TA_BAND / 4)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if (!FR || nt < nth) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int rl = warp * 16 + tf32::c_row(0, e);
            Rs[rl * TA_RLD + rl - tf32::c_col(nt, e) + 63] = s[0][nt][e];
          }
        }
      __syncwarp();
      // dQ += dS' K (dS'[i, j] read back from its skewed slot) + dR E_win (only the warp's band of window columns is non-zero; with nl live keys only its n-tiles from ntf on)
      tf32::warp_gemm_fr<FR, 1, 2>(dq, 8, 0, nth, 0, 2, [&](int r, int k) { return Rs[(warp * 16 + r) * TA_RLD + warp * 16 + r - k + 63]; },
                                   [&](int k, int c) { return Ks[k * TA_LD + c]; });
      tf32::warp_gemm_fr<FR, 1, 2>(dq, TA_BAND / 8, ntf, TA_BAND / 8, 0, 2, [&](int r, int k) { return Rs[(warp * 16 + r) * TA_RLD + warp * 16 + k]; },
                                   [&](int k, int c) { return Es[(warp * 16 + k) * TA_LD + c]; });
    };
    for (int j0 = 0; j0 < n; j0 += TA_B) {
      __syncthreads();                                  // previous tile's dE GEMM has read Rs / Qs; Ks / Vs / Es are free
      ta_load_tile(Ks, qkv, base, sq.pos_stride, TA_ROW, 64 + h * TA_D, j0, n);
      ta_load_tile(Vs, qkv, base, sq.pos_stride, TA_ROW, 128 + h * TA_D, j0, n);
      const int dbase = i0 - j0 - 63;
      ta_load_window(Es, E, dbase);
      __syncthreads();
      const int nl = ta_live(n, j0);
      if (warp_live) {
        if (nl == TA_B) tile_rows(std::false_type{}, j0); else tile_rows(std::true_type{}, j0);
      } else {                                          // dead rows: their dR band must read as zero in the dE GEMM below (it may hold the previous item's values)
        for (int idx = threadIdx.x & 31; idx < 16 * (TA_BAND / 4); idx += 32)
          *reinterpret_cast<float4*>(Rs + (warp * 16 + idx / (TA_BAND / 4)) * TA_RLD + warp * 16 + (idx % (TA_BAND / 4)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();                                  // all 64 rows of dR are in place
      {
        // dE_win[w] = sum_i dR[i, w] q_i for this warp's 32 window rows; dR[i, w] != 0 only for w - 63 <= i <= w, i < nr and w >= i + 64 - nl
        const int k_lo = warp == 3 ? 32 : 0, k_n = (warp == 0 || warp == 3) ? 4 : 8;
        int k_hi = (nr - k_lo + 7) >> 3;                // k-steps that still hold live query rows
        k_hi = k_hi < 0 ? 0 : (k_hi > k_n ? k_n : k_hi);
        if (warp * 32 + 31 >= TA_B - nl && k_hi > 0) {
          float de[2][2][4];
          tf32::zero(de);
          tf32::warp_gemm_range<2, 2>(de, 0, k_hi, 0, 2, [&](int r, int k) { return Rs[(k_lo + k) * TA_RLD + warp * 32 + r]; }, [&](int k, int c) { return Qs[(k_lo + k) * TA_LD + c]; });
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int w = warp * 32 + tf32::c_row(mt, e);
                if (w < 127) {
                  int d = dbase + w;
                  d = d < -emax ? -emax : (d > emax ? emax : d);          // |d| <= n - 1 wherever dR is non-zero; the clamp only acts at +-512
                  atomicAdd(dEt + (d + emax) * TA_D + tf32::c_col(nt, e), de[mt][nt][e]);
                }
              }
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const int i = i0 + warp * 16 + tf32::c_row(0, e);
        if (i < n)
          *reinterpret_cast<float2*>(dqkv + (base + (long long)i * sq.pos_stride) * TA_ROW + h * TA_D + tf32::c_col(nt, e)) = make_float2(dq[0][nt][e], dq[0][nt][e + 1]);
      }
  }
  __syncthreads();
  float* dst = de_partial + (long long)blockIdx.x * (TA_EROWS * TA_D) + (TA_MAXPOS - emax) * TA_D;      // rows -emax .. +emax of the full table
  for (int i = threadIdx.x; i < trows * TA_D; i += 128) dst[i] = dEt[i];
}

// dE[r] = sum over CTAs of their partial tables (rows the sequences cannot address are zero)
__global__ void __launch_bounds__(256) attention_de_finish_kernel(const float* __restrict__ partial, int rows, int emax, float* __restrict__ dE) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= TA_EROWS * TA_D) return;
  const int r = c / TA_D - TA_MAXPOS;
  float s = 0.f;
  if (r >= -emax && r <= emax)
    for (int k = 0; k < rows; ++k) s += partial[(long long)k * (TA_EROWS * TA_D) + c];
  dE[c] = s;
}

// ------------------------------------------------------------------------------------------------------------------------------------
// backward, key-major: dK, dV (into dqkv[:, 64:192])
// ------------------------------------------------------------------------------------------------------------------------------------
// The E window shares the memory of the transposed P / dS tile: it is dead once every warp has its R rows (second block barrier of a tile) and Ts is
// first written after that barrier; 72 KB instead of 82 KB lets three CTAs (12 warps) share an SM instead of two.
constexpr int TAK_SMEM = (4 * TA_B * TA_LD + TA_B * TA_RLD + TA_B * TA_PLD + 2 * TA_B) * 4;
static_assert(TA_B * TA_PLD >= 128 * TA_LD, "the E window must fit inside the transposed tile");

__global__ void __launch_bounds__(128) attention_bwd_k_kernel(const float* __restrict__ qkv, const float* __restrict__ E, const SebSeq sq, int nkt,
                                                             const float* __restrict__ lse, const float* __restrict__ Dg, const float* __restrict__ d_o,
                                                             float* __restrict__ dqkv) {
  extern __shared__ __align__(16) float ta_sm[];
  float* Qs = ta_sm; float* dOs = Qs + TA_B * TA_LD; float* Ks = dOs + TA_B * TA_LD; float* Vs = Ks + TA_B * TA_LD;
  float* Rs = Vs + TA_B * TA_LD; float* Ts = Rs + TA_B * TA_RLD; float* Es = Ts; float* Ls = Ts + TA_B * TA_PLD; float* Ds = Ls + TA_B;
  const int warp = threadIdx.x >> 5;
  const int kt = blockIdx.x % nkt, sh = blockIdx.x / nkt, h = sh & 3, seq = sh >> 2;
  const int n = sq.n, j0 = kt * TA_B;
  const long long base = ta_seq_base(sq, seq);
  ta_load_tile(Ks, qkv, base, sq.pos_stride, TA_ROW, 64 + h * TA_D, j0, n);
  ta_load_tile(Vs, qkv, base, sq.pos_stride, TA_ROW, 128 + h * TA_D, j0, n);
  float dk[1][2][4], dv[1][2][4];
  tf32::zero(dk);
  tf32::zero(dv);
  const int nlk = ta_live(n, j0);                       // live keys of this CTA's tile
  const bool key_live = warp * 16 < nlk;                // a warp whose 16 keys all lie past the sequence only keeps the barriers (and its share of the R rows)
  // the key-owned part of one query tile; FR: fewer than 64 live queries (n = 64 k + 1 leaves ONE in the last tile) -- the dead n-tiles / k-steps are skipped
  auto tile_keys = [&](auto fr_tag, int i0) {
    constexpr bool FR = decltype(fr_tag)::value;
    const int nth = (ta_live(n, i0) + 7) >> 3;
    // S^T[j, i] = k_j . q_i;  dP^T[j, i] = v_j . dO_i   (rows = this warp's 16 keys, columns = the 64 queries)
    float st[1][8][4], dpt[1][8][4];
    tf32::zero(st);
    tf32::zero(dpt);
    tf32::warp_gemm_fr<FR, 1, 8>(st, 2, 0, 2, 0, nth, [&](int r, int k) { return Ks[(warp * 16 + r) * TA_LD + k]; }, [&](int k, int c) { return Qs[c * TA_LD + k]; });
    tf32::warp_gemm_fr<FR, 1, 8>(dpt, 2, 0, 2, 0, nth, [&](int r, int k) { return Vs[(warp * 16 + r) * TA_LD + k]; }, [&](int k, int c) { return dOs[c * TA_LD + k]; });
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (!FR || nt < nth) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int jl = warp * 16 + tf32::c_row(0, e), il = tf32::c_col(nt, e);
          const float v = TA_SCALE * (st[0][nt][e] + Rs[il * TA_RLD + il - jl + 63]);
          const float p = (j0 + jl < n && i0 + il < n) ? expf(v - Ls[il]) : 0.f;
          st[0][nt][e] = p;
          dpt[0][nt][e] = TA_SCALE * p * (dpt[0][nt][e] - Ds[il]);
          Ts[jl * TA_PLD + il] = p;
        }
      }
    __syncwarp();
    tf32::warp_gemm_fr<FR, 1, 2>(dv, 8, 0, nth, 0, 2, [&](int r, int k) { return Ts[(warp * 16 + r) * TA_PLD + k]; }, [&](int k, int c) { return dOs[k * TA_LD + c]; });
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (!FR || nt < nth) {
#pragma unroll
        for (int e = 0; e < 4; ++e) Ts[(warp * 16 + tf32::c_row(0, e)) * TA_PLD + tf32::c_col(nt, e)] = dpt[0][nt][e];
      }
    __syncwarp();
    tf32::warp_gemm_fr<FR, 1, 2>(dk, 8, 0, nth, 0, 2, [&](int r, int k) { return Ts[(warp * 16 + r) * TA_PLD + k]; }, [&](int k, int c) { return Qs[k * TA_LD + c]; });
  };
  for (int i0 = 0; i0 < n; i0 += TA_B) {
    __syncthreads();
    ta_load_tile(Qs, qkv, base, sq.pos_stride, TA_ROW, h * TA_D, i0, n);
    ta_load_tile(dOs, d_o, base, sq.pos_stride, 64, h * TA_D, i0, n);
    ta_load_window(Es, E, i0 - j0 - 63);
    if (threadIdx.x < TA_B) {
      const int i = i0 + threadIdx.x;
      const long long tok = base + (long long)(i < n ? i : 0) * sq.pos_stride;
      Ls[threadIdx.x] = i < n ? lse[tok * 4 + h] : 0.f;
      Ds[threadIdx.x] = i < n ? Dg[tok * 4 + h] : 0.f;
    }
    __syncthreads();
    const int nrq = ta_live(n, i0);                     // live query rows of this tile
    if (warp * 16 < nrq) {                              // warp w: R rows of queries 16 w .. + 15 (skipped when they all lie past the sequence)
      if (nlk == TA_B) ta_rel_rows<false>(Qs, Es, Rs, warp, 0); else ta_rel_rows<true>(Qs, Es, Rs, warp, (TA_B - nlk) >> 3);
    }
    __syncthreads();                                    // the transposed tile below reads every query row
    if (key_live) {
      if (nrq == TA_B) tile_keys(std::false_type{}, i0); else tile_keys(std::true_type{}, i0);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; e += 2) {
      const int j = j0 + warp * 16 + tf32::c_row(0, e);
      if (j < n) {
        float* row = dqkv + (base + (long long)j * sq.pos_stride) * TA_ROW + h * TA_D + tf32::c_col(nt, e);
        *reinterpret_cast<float2*>(row + 64) = make_float2(dk[0][nt][e], dk[0][nt][e + 1]);
        *reinterpret_cast<float2*>(row + 128) = make_float2(dv[0][nt][e], dv[0][nt][e + 1]);
      }
    }
}

static int ta_check(const void* qkv, const float* E, const SebSeq* seq) {
  SEB_REQUIRE(qkv && E && seq && aligned16(qkv) && aligned16(E), SEB_EINVAL, "attention (train): null / unaligned argument");
  SEB_REQUIRE(seq->nseq > 0 && seq->n > 0 && seq->inner > 0 && seq->nseq <= (1 << 26), SEB_EINVAL, "attention (train): bad sequence descriptor");
  return 0;
}

}  // namespace seb

using namespace seb;

// fp32-grade forward for the training step: qkv fp32 [tokens, 192] (unscaled) -> out [tokens, 64], lse [tokens, 4] (natural log)
extern "C" int seb200_attention_train_fwd(const float* qkv, const float* rel_pos_emb, const SebSeq* seq, float* out, float* lse, void* stream) {
  if (int rc = ta_check(qkv, rel_pos_emb, seq)) return rc;
  SEB_REQUIRE(out && lse && aligned16(out), SEB_EINVAL, "attention_train_fwd: null output");
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(attention_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAF_SMEM);
    if (e != cudaSuccess) { set_error("attention_train_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  const int nqt = (seq->n + TA_B - 1) / TA_B;
  const long long nb = (long long)seq->nseq * 4 * nqt;
  SEB_REQUIRE(nb < 2147483647LL, SEB_EINVAL, "attention_train_fwd: grid too large");
  attention_train_fwd_kernel<<<(unsigned)nb, 128, TAF_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, rel_pos_emb, *seq, nqt, out, lse);
  SEB_CHECK_LAUNCH("attention_train_fwd_kernel");
  return 0;
}

constexpr int TAQ_CTAS = 2 * 148;
extern "C" long long seb200_attention_bwd_workspace_floats(long long tokens) { return tokens * 4 + (long long)TAQ_CTAS * TA_EROWS * TA_D; }

// backward: dqkv [tokens, 192] (every element written), drel [1025, 16]
extern "C" int seb200_attention_bwd(const float* qkv, const float* rel_pos_emb, const SebSeq* seq, long long tokens, const float* out, const float* lse,
                                    const float* dout, float* dqkv, float* drel, float* workspace, long long workspace_floats, void* stream) {
  if (int rc = ta_check(qkv, rel_pos_emb, seq)) return rc;
  SEB_REQUIRE(out && lse && dout && dqkv && drel && workspace && aligned16(out) && aligned16(dout) && aligned16(dqkv) && aligned16(workspace), SEB_EINVAL,
              "attention_bwd: null / unaligned argument");
  SEB_REQUIRE(tokens >= (long long)seq->nseq * seq->n && workspace_floats >= seb200_attention_bwd_workspace_floats(tokens), SEB_EINVAL, "attention_bwd: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, taq_smem(TA_MAXPOS + 1));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_bwd_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAK_SMEM);
    if (e != cudaSuccess) { set_error("attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  float* D = workspace;
  float* de_partial = workspace + tokens * 4;
  {
    long long g = (tokens * 4 + 255) / 256; if (g > 148 * 8) g = 148 * 8;
    attention_rowdot_kernel<<<(unsigned)g, 256, 0, st>>>(out, dout, tokens * 4, D);
    SEB_CHECK_LAUNCH("attention_rowdot_kernel");
  }
  const int nt = (seq->n + TA_B - 1) / TA_B;
  const long long nitems = (long long)seq->nseq * 4 * nt;
  SEB_REQUIRE(nitems < 2147483647LL, SEB_EINVAL, "attention_bwd: grid too large");
  const int smem_q = taq_smem(seq->n);
  const int per_sm = (227 * 1024) / (smem_q + 1024) >= 2 ? 2 : 1;            // two persistent CTAs per SM when the table is small enough (n <= ~400)
  const int nq = (int)(nitems < per_sm * 148 ? nitems : per_sm * 148);
  attention_bwd_q_kernel<<<nq, 128, smem_q, st>>>(qkv, rel_pos_emb, *seq, nt, nitems, lse, D, dout, dqkv, de_partial);
  SEB_CHECK_LAUNCH("attention_bwd_q_kernel");
  const int emax = seq->n - 1 < TA_MAXPOS ? seq->n - 1 : TA_MAXPOS;
  attention_de_finish_kernel<<<(TA_EROWS * TA_D + 255) / 256, 256, 0, st>>>(de_partial, nq, emax, drel);
  SEB_CHECK_LAUNCH("attention_de_finish_kernel");
  attention_bwd_k_kernel<<<(unsigned)nitems, 128, TAK_SMEM, st>>>(qkv, rel_pos_emb, *seq, nt, lse, D, dout, dqkv);
  SEB_CHECK_LAUNCH("attention_bwd_k_kernel");
  return 0;
}
