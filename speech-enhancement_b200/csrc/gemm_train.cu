// gemm_train.cu -- the GEMM engine's instantiations for the training step (SEB_ENGINE_TCGEN05_F32): the generic
// tcgen05 main loop of gemm_engine.cuh with THREE bf16 planes per operand (hi | mid | lo, six products: every term down to 2^-24, as the
// DFT / iDFT use it), fp32 activations split on the fly by the producer warps.  Forward GEMMs of the train-mode generator, their dgrad
// (the same contraction with the transposed weight image, or the adjoint-conv loader) -- the wgrad contractions reduce over the pixels /
// tokens instead and live in wgrad.cu.
#include "gemm_engine.cuh"

namespace seb {

template <int NT, int STAGES, int LK, int EK>
static int launch_tc3(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  constexpr int SMEM = tc_smem_bytes<NT, STAGES, 3>();
  static_assert(SMEM <= 227 * 1024, "stage ring exceeds the shared memory of an SM");
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<NT, STAGES, LK, EK, 8, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("gemm train: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  SEB_REQUIRE(s->w_tc && s->tc_planes == 3 && s->tc_ntile == NT && s->tc_ntiles >= 1 && s->tc_ntile * s->tc_ntiles >= s->N && aligned16(s->w_tc), SEB_EINVAL,
              "gemm train: weight image has n-tile %d x %d / %d planes, kernel wants %d / 3 covering N=%d", s->tc_ntile, s->tc_ntiles, s->tc_planes, NT, s->N);
  dim3 grid((g.M + BM - 1) / BM, s->tc_ntiles);
  gemm_tc_kernel<NT, STAGES, LK, EK, 8, 1, 3><<<grid, 10 * 32, SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc));
  SEB_CHECK_LAUNCH("gemm_tc_kernel<3 planes>");
  return 0;
}

int launch_train_tc(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  const int nt = s->tc_ntile;
  switch (s->loader * 16 + s->epilogue) {
    case SEB_LOAD_ROWS * 16 + SEB_EPI_BIAS:
      if (nt == 64) return launch_tc3<64, 2, SEB_LOAD_ROWS, SEB_EPI_BIAS>(s, g, st);
      if (nt == 128) return launch_tc3<128, 1, SEB_LOAD_ROWS, SEB_EPI_BIAS>(s, g, st);
      if (nt == 256) return launch_tc3<256, 1, SEB_LOAD_ROWS, SEB_EPI_BIAS>(s, g, st);
      break;
    case SEB_LOAD_ROWS * 16 + SEB_EPI_RESID:    if (nt == 64) return launch_tc3<64, 2, SEB_LOAD_ROWS, SEB_EPI_RESID>(s, g, st); break;
    case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_BIAS:
      if (nt == 192) return launch_tc3<192, 1, SEB_LOAD_ROWS_LN, SEB_EPI_BIAS>(s, g, st);
      if (nt == 256) return launch_tc3<256, 1, SEB_LOAD_ROWS_LN, SEB_EPI_BIAS>(s, g, st);
      break;
    case SEB_LOAD_CONV * 16 + SEB_EPI_BIAS:     if (nt == 64) return launch_tc3<64, 2, SEB_LOAD_CONV, SEB_EPI_BIAS>(s, g, st); break;
    case SEB_LOAD_CONV * 16 + SEB_EPI_SUBPIXEL: if (nt == 128) return launch_tc3<128, 1, SEB_LOAD_CONV, SEB_EPI_SUBPIXEL>(s, g, st); break;
    case SEB_LOAD_CONV_ADJ * 16 + SEB_EPI_BIAS: if (nt == 64) return launch_tc3<64, 2, SEB_LOAD_CONV_ADJ, SEB_EPI_BIAS>(s, g, st); break;
    case SEB_LOAD_CONV_ADJ * 16 + SEB_EPI_RESID: if (nt == 64) return launch_tc3<64, 2, SEB_LOAD_CONV_ADJ, SEB_EPI_RESID>(s, g, st); break;
    default: break;
  }
  return -100;
}

}  // namespace seb
