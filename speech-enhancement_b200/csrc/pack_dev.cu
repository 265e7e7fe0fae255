// pack_dev.cu -- the weight packer of pack.cu as a KERNEL: in training the parameters change every optimizer step
// (/root/reference/core/function.py:277), so the GEMM engine's weight images are rebuilt on the device from the live fp32 parameters
// instead of on the host.  Same images, bit for bit (tests compare them with seb200_pack_weights):
//   w_tc   [n-tile][k-chunk][plane][tc_ntile rows][64 bf16], 128-byte rows in the UMMA K-major SWIZZLE_128B form, planes hi | lo or hi | mid | lo
//   w_simt fp32 [kp][npad]
// The logical matrix W[n, k] is read through a two-level index map  w[n * sn + (k / n1) * s0 + (k % n1) * s1]  (k < K), which covers
//   * nn.Linear / pointwise Conv1d weights [N, K]:                    n1 = K, sn = K, s1 = 1
//   * their transposes (dgrad: dX = dY . W):                            n1 = K', sn = 1, s1 = N'
//   * Conv2d weights [Cout, Cin, kt, kf] in the engine's K order (tap, cin):   n1 = Cin, sn = Cin * taps, s0 = 1, s1 = taps
//   * the adjoint conv of one 64-channel input slot j (dgrad):  W'[ci, (tap, co)] = w[co, 64 j + ci, tap]:  n1 = Cout, sn = taps, s0 = 1, s1 = Cin * taps
#include "common.cuh"

namespace seb {

__device__ __forceinline__ uint16_t bf16_rne_dev(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_f32_dev(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }

// one packing job (device pointers): the arguments of seb200_pack_weights_device plus the derived sizes
struct PackJob {
  const float* w; uint16_t* img; float* w_simt;
  long long sn, s0, s1;
  int N, K, n1, tc_ntile, planes, kp, ntiles, npad;
};
constexpr int PACK_BATCH = 24;                       // jobs per launch: the descriptor array travels as a kernel parameter (< 4 KB)
struct PackBatch { PackJob job[PACK_BATCH]; int first_block[PACK_BATCH + 1]; int njobs; };

__device__ __forceinline__ void pack_item(const PackJob& J, long long idx) {
  const float* __restrict__ w = J.w;
  const int N = J.N, K = J.K, n1 = J.n1, tc_ntile = J.tc_ntile, planes = J.planes, kp = J.kp, npad = J.npad;
  const long long sn = J.sn, s0 = J.s0, s1 = J.s1;
  const int nkc = kp / 64;
  const long long tc_items = J.img ? (long long)J.ntiles * nkc * tc_ntile * 8 : 0;       // one item = one 16-byte chunk of one (padded) row
  if (idx < tc_items) {
    const int c = (int)(idx & 7);
    long long t = idx >> 3;
    const int r = (int)(t % tc_ntile); t /= tc_ntile;
    const int kc = (int)(t % nkc);
    const int j = (int)(t / nkc);
    const int n = j * tc_ntile + r;
    uint32_t wd[3][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * 64 + c * 8 + i;
      const float x = (n < N && k < K) ? w[(long long)n * sn + (long long)(k / n1) * s0 + (long long)(k % n1) * s1] : 0.f;
      const uint16_t h = bf16_rne_dev(x);
      const float r1 = x - bf16_f32_dev(h);
      const uint16_t m = bf16_rne_dev(r1);
      const uint16_t l = bf16_rne_dev(r1 - bf16_f32_dev(m));
      const int sh = (i & 1) * 16;
      if ((i & 1) == 0) { wd[0][i >> 1] = 0u; wd[1][i >> 1] = 0u; wd[2][i >> 1] = 0u; }
      wd[0][i >> 1] |= (uint32_t)h << sh;
      wd[1][i >> 1] |= (uint32_t)m << sh;
      wd[2][i >> 1] |= (uint32_t)l << sh;
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      if (p < planes) {
        uint16_t* blk = J.img + ((((long long)j * nkc + kc) * planes + p) * tc_ntile) * 64 + r * 64 + ((c ^ (r & 7)) << 3);
        *reinterpret_cast<uint4*>(blk) = make_uint4(wd[p][0], wd[p][1], wd[p][2], wd[p][3]);
      }
    }
    return;
  }
  if (J.w_simt) {
    const long long e = idx - tc_items;
    if (e < (long long)kp * npad) {
      const int k = (int)(e / npad), n = (int)(e - (long long)k * npad);
      J.w_simt[e] = (k < K && n < N) ? w[(long long)n * sn + (long long)(k / n1) * s0 + (long long)(k % n1) * s1] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int N, int K, int n1, long long sn, long long s0, long long s1,
                                                           int tc_ntile, int planes, int kp, int ntiles, int npad,
                                                           uint16_t* __restrict__ img, float* __restrict__ w_simt) {
  PackJob J;
  J.w = w; J.img = img; J.w_simt = w_simt; J.sn = sn; J.s0 = s0; J.s1 = s1;
  J.N = N; J.K = K; J.n1 = n1; J.tc_ntile = tc_ntile; J.planes = planes; J.kp = kp; J.ntiles = ntiles; J.npad = npad;
  pack_item(J, (long long)blockIdx.x * 256 + threadIdx.x);
}

// up to PACK_BATCH jobs in one launch: block b belongs to the job j with first_block[j] <= b < first_block[j + 1] (a training step packs 176 images;
// one launch each was 1.1 ms of 3 - 6 us kernels)
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const __grid_constant__ PackBatch B) {
  int j = 0;
  while (j + 1 < B.njobs && (int)blockIdx.x >= B.first_block[j + 1]) ++j;
  pack_item(B.job[j], (long long)(blockIdx.x - B.first_block[j]) * 256 + threadIdx.x);
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_pack_weights_device(const float* w, int N, int K, int n1, long long sn, long long s0, long long s1, int tc_ntile, int planes,
                                          void* w_tc, float* w_simt, void* stream) {
  SEB_REQUIRE(w && (w_tc || w_simt) && n1 > 0, SEB_EINVAL, "pack_weights_device: null pointer");
  long long tcb = 0, sf = 0;
  int kp = 0, ntiles = 0, npad = 0;
  const int rc = seb200_packed_weight_sizes(N, K, tc_ntile, planes, &tcb, &sf, &kp, &ntiles, &npad);
  if (rc) return rc;
  SEB_REQUIRE(!w_tc || aligned16(w_tc), SEB_EALIGN, "pack_weights_device: image not 16-byte aligned");
  const long long items = (w_tc ? (long long)ntiles * (kp / 64) * tc_ntile * 8 : 0) + (w_simt ? (long long)kp * npad : 0);
  pack_weights_kernel<<<(unsigned)((items + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, N, K, n1, sn, s0, s1, tc_ntile, planes, kp, ntiles, npad, reinterpret_cast<uint16_t*>(w_tc), w_simt);
  SEB_CHECK_LAUNCH("pack_weights_kernel");
  return 0;
}

// The same for `njobs` images at once (host array of job descriptors with DEVICE pointers): ceil(njobs / 24) launches instead of njobs.
extern "C" int seb200_pack_weights_device_batch(const SebPackJob* jobs, int njobs, void* stream) {
  SEB_REQUIRE(jobs && njobs > 0, SEB_EINVAL, "pack_weights_device_batch: no jobs");
  for (int base = 0; base < njobs; base += PACK_BATCH) {
    PackBatch B;
    B.njobs = njobs - base < PACK_BATCH ? njobs - base : PACK_BATCH;
    int blocks = 0;
    for (int i = 0; i < B.njobs; ++i) {
      const SebPackJob& s = jobs[base + i];
      SEB_REQUIRE(s.w && (s.w_tc || s.w_simt) && s.n1 > 0, SEB_EINVAL, "pack_weights_device_batch: job %d: null pointer", base + i);
      SEB_REQUIRE(!s.w_tc || aligned16(s.w_tc), SEB_EALIGN, "pack_weights_device_batch: job %d: image not 16-byte aligned", base + i);
      long long tcb = 0, sf = 0;
      int kp = 0, ntiles = 0, npad = 0;
      const int rc = seb200_packed_weight_sizes(s.N, s.K, s.tc_ntile, s.planes, &tcb, &sf, &kp, &ntiles, &npad);
      if (rc) return rc;
      PackJob& J = B.job[i];
      J.w = s.w; J.img = reinterpret_cast<uint16_t*>(s.w_tc); J.w_simt = s.w_simt; J.sn = s.sn; J.s0 = s.s0; J.s1 = s.s1;
      J.N = s.N; J.K = s.K; J.n1 = s.n1; J.tc_ntile = s.tc_ntile; J.planes = s.planes; J.kp = kp; J.ntiles = ntiles; J.npad = npad;
      const long long items = (s.w_tc ? (long long)ntiles * (kp / 64) * s.tc_ntile * 8 : 0) + (s.w_simt ? (long long)kp * npad : 0);
      B.first_block[i] = blocks;
      blocks += (int)((items + 255) / 256);
    }
    B.first_block[B.njobs] = blocks;
    pack_weights_batch_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(B);
    SEB_CHECK_LAUNCH("pack_weights_batch_kernel");
  }
  return 0;
}
