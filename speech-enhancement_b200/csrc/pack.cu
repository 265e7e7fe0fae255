// pack.cu -- host-side helpers of the C ABI: weight packing for the GEMM engine and workspace sizing (SURVEY 8b:
// `seb200_pack_weights`, `seb200_workspace_bytes`).  Pure host code (no launches): a caller that is not the shipped Python host
// packs a layer's fp32 weight matrix here, copies the two images to the device and passes them in SebGemm / SebFfn.
//
// Images of one logical W [N, K] fp32 (row-major), see seb200.h:
//   w_tc   : K padded to kp = ceil(K / 64) * 64, N padded to ntiles * tc_ntile.  W = hi + lo (+ mid) in bf16, round to nearest even:
//            hi = bf16(W), lo = bf16(W - hi)   [3 planes: mid = bf16(W - hi), lo = bf16(W - hi - mid)].
//            Layout [n-tile][k-chunk][plane][tc_ntile rows][64 bf16]: every (tile, chunk, plane) block is tc_ntile rows of 128 bytes
//            in the UMMA K-major SWIZZLE_128B canonical form -- the 16-byte chunk c of row r sits at byte r * 128 + ((c ^ (r & 7)) << 4)
//            -- so one cp.async.bulk per pipeline stage lands MMA-ready operands.
//   w_simt : fp32 [kp][npad], npad = ceil(N / 64) * 64: W transposed (K-major rows), zero padded (the fp32 FFMA main loop).
#include <string.h>
#include "common.cuh"

namespace seb {

static inline uint16_t bf16_rne(float f) {            // round to nearest even, like torch's float -> bfloat16
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);      // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_packed_weight_sizes(int N, int K, int tc_ntile, int planes, long long* tc_bytes, long long* simt_floats,
                                          int* k_padded, int* tc_ntiles, int* simt_npad) {
  SEB_REQUIRE(N > 0 && K > 0 && tc_ntile >= 16 && tc_ntile <= 256 && tc_ntile % 16 == 0 && (planes == 2 || planes == 3), SEB_EINVAL,
              "packed_weight_sizes: bad arguments N=%d K=%d n-tile=%d planes=%d", N, K, tc_ntile, planes);
  const int kp = (K + 63) / 64 * 64, ntiles = (N + tc_ntile - 1) / tc_ntile, npad = (N + 63) / 64 * 64;
  if (tc_bytes) *tc_bytes = (long long)ntiles * (kp / 64) * planes * tc_ntile * 128;
  if (simt_floats) *simt_floats = (long long)kp * npad;
  if (k_padded) *k_padded = kp;
  if (tc_ntiles) *tc_ntiles = ntiles;
  if (simt_npad) *simt_npad = npad;
  return 0;
}

extern "C" int seb200_pack_weights(const float* w, int N, int K, int tc_ntile, int planes, void* w_tc, float* w_simt) {
  SEB_REQUIRE(w && (w_tc || w_simt), SEB_EINVAL, "pack_weights: null pointer");
  long long tcb = 0, sf = 0;
  int kp = 0, ntiles = 0, npad = 0;
  const int rc = seb200_packed_weight_sizes(N, K, tc_ntile, planes, &tcb, &sf, &kp, &ntiles, &npad);
  if (rc) return rc;
  if (w_tc) {
    uint16_t* img = reinterpret_cast<uint16_t*>(w_tc);
    const int nkc = kp / 64;
    for (int j = 0; j < ntiles; ++j)
      for (int kc = 0; kc < nkc; ++kc)
        for (int r = 0; r < tc_ntile; ++r) {
          const int n = j * tc_ntile + r;
          for (int c = 0; c < 8; ++c) {                      // 16-byte chunk = 8 consecutive k
            uint16_t pl[3][8];
            for (int i = 0; i < 8; ++i) {
              const int k = kc * 64 + c * 8 + i;
              const float x = (n < N && k < K) ? w[(long long)n * K + k] : 0.f;
              const uint16_t h = bf16_rne(x);
              const float r1 = x - bf16_to_f32(h);
              pl[0][i] = h;
              if (planes == 2) {
                pl[1][i] = bf16_rne(r1);
              } else {
                const uint16_t m = bf16_rne(r1);
                pl[1][i] = m;
                pl[2][i] = bf16_rne(r1 - bf16_to_f32(m));
              }
            }
            for (int p = 0; p < planes; ++p) {
              uint16_t* blk = img + ((((long long)j * nkc + kc) * planes + p) * tc_ntile) * 64;
              memcpy(blk + r * 64 + ((c ^ (r & 7)) << 3), pl[p], 16);
            }
          }
        }
  }
  if (w_simt) {
    for (int k = 0; k < kp; ++k)
      for (int n = 0; n < npad; ++n) w_simt[(long long)k * npad + n] = (k < K && n < N) ? w[(long long)n * K + k] : 0.f;
  }
  return 0;
}

// Activation workspace of one forward as the shipped host allocates it (generator.TSCNet.workspace): kind 0 = GAN generator
// (models/generator.py), kind 1 = diffusion variant (+ the conditioning encoder's output).  B utterances, T frames, F bins (odd).
extern "C" long long seb200_workspace_bytes(int kind, int B, int T, int F) {
  if (B <= 0 || T <= 0 || F <= 0 || F % 2 == 0 || (kind != 0 && kind != 1)) {
    set_error("workspace_bytes: bad arguments kind=%d B=%d T=%d F=%d", kind, B, T, F);
    return -1;
  }
  const long long Fh = (F - 1) / 2 + 1, P = (long long)B * T * F, Ph = (long long)B * T * Fh, BT = (long long)B * T;
  long long b = 0;
  b += 6 * P * 64 * 4;                        // encoder @F: conv_1 / dense outputs (5) + raw conv output
  b += 5 * Ph * 64 * 4;                       // decoders @F': dense outputs (4) + raw
  b += BT * 2 * Fh * 64 * 4;                  // sub-pixel output @2F'
  b += 4 * Ph * 64 * 4;                       // xs (pre-split TSCB output), x, y, o
  b += Ph * 256 * 4 + Ph * 192 * 4;           // h (SIMT feed-forward hidden), q|k|v
  b += 2 * Ph * 128 * 4;                      // u (GLU output), v (depthwise output)
  b += BT * F * 4 + 2 * BT * F * 2 * 4;       // mask_raw, cplx, est
  b += (long long)B * 64 * 2 * 4 + (long long)B * 2 * 4;                  // InstanceNorm statistics (64-channel, 1-channel)
  b += (seb200_inorm_workspace_bytes(B, (long long)T * 2 * Fh, 64) + 7) / 8 * 8;
  if (kind == 1) b += Ph * 64 * 4;            // cond
  return b;
}
