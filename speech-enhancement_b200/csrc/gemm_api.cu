// gemm_api.cu -- seb200_gemm: argument checking and dispatch onto the instantiated engine kernels.
#include "gemm_engine.cuh"
#include <stdlib.h>

namespace seb {

int launch_conv_persist(const SebGemm* s, const GemmArgs& g, cudaStream_t st);   // conv_persist.cu
int conv_persist_max_chunks();
int launch_conv_y3(const SebGemm* s, const GemmArgs& g, cudaStream_t st);        // conv_y3.cu
int conv_y3_enabled();
int launch_tok_gemm(const SebGemm* s, const GemmArgs& g, cudaStream_t st);       // tok_gemm.cu (-100: not a persistent token GEMM)
int launch_train_tc(const SebGemm* s, const GemmArgs& g, cudaStream_t st);       // gemm_train.cu: three-plane tcgen05 instantiations of the training step

static GemmArgs to_args(const SebGemm* s) {
  GemmArgs g;
  for (int i = 0; i < 4; ++i) g.a[i] = s->a[i];
  g.lda = s->lda; g.ln_g = s->ln_gamma; g.ln_b = s->ln_beta;
  g.M = s->M; g.N = s->N; g.K = s->K;
  g.B = s->B; g.T = s->T; g.Fin = s->Fin; g.Fout = s->Fout;
  g.taps_t = s->taps_t; g.dil = s->dil; g.stride_f = s->stride_f; g.nslots = s->nslots;
  g.bias = s->bias; g.out = s->out; g.ldo = s->ldo; g.resid = s->resid; g.ldr = s->ldr; g.alpha = s->alpha;
  return g;
}

template <int LK, int EK>
static int launch_simt(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_simt_kernel<LK, EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SIMT_SMEM);
    if (e != cudaSuccess) { set_error("gemm simt: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  SEB_REQUIRE(s->w_simt && s->simt_npad % 64 == 0 && s->simt_npad >= s->N, SEB_EINVAL, "gemm simt: bad weight image");
  dim3 grid((g.M + BM - 1) / BM, s->simt_npad / 64);
  gemm_simt_kernel<LK, EK><<<grid, 256, SIMT_SMEM, st>>>(g, s->w_simt, s->simt_npad);
  SEB_CHECK_LAUNCH("gemm_simt_kernel");
  return 0;
}

template <int NT, int STAGES, int LK, int EK, int PW = 8, int MINB = 2, int NPL = 2>
static int launch_tc(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  constexpr int SMEM = tc_smem_bytes<NT, STAGES, NPL>();
  SEB_REQUIRE(s->tc_planes == NPL, SEB_EINVAL, "gemm tc: weight image has %d planes, kernel wants %d", s->tc_planes, NPL);
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<NT, STAGES, LK, EK, PW, MINB, NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("gemm tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  SEB_REQUIRE(s->w_tc && s->tc_ntile == NT && s->tc_ntiles >= 1 && s->tc_ntile * s->tc_ntiles >= s->N, SEB_EINVAL,
              "gemm tc: weight image has n-tile %d x %d, kernel wants %d covering N=%d", s->tc_ntile, s->tc_ntiles, NT, s->N);
  SEB_REQUIRE(aligned16(s->w_tc), SEB_EALIGN, "gemm tc: weight image not 16-byte aligned");
  dim3 grid((g.M + BM - 1) / BM, s->tc_ntiles);
  gemm_tc_kernel<NT, STAGES, LK, EK, PW, MINB, NPL><<<grid, (PW + 2) * 32, SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc));
  SEB_CHECK_LAUNCH("gemm_tc_kernel");
  return 0;
}

template <int NT, int STAGES, int EK>
static int launch_conv_split(const SebGemm* s, const GemmArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_done;
  constexpr int SMEM = tc_smem_bytes<NT, STAGES>();
  if (!attr_done.done()) {
    cudaError_t e = cudaFuncSetAttribute(conv_split_tc_kernel<NT, STAGES, EK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("conv split: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.set();
  }
  SEB_REQUIRE(s->w_tc && s->tc_ntile == NT && s->tc_ntiles >= 1 && s->tc_ntile * s->tc_ntiles >= s->N && aligned16(s->w_tc), SEB_EINVAL,
              "conv split: weight image has n-tile %d x %d, kernel wants %d covering N=%d", s->tc_ntile, s->tc_ntiles, NT, s->N);
  SEB_REQUIRE(s->T < 32768 && s->Fout < 65536 && (long long)s->B * s->T * s->Fin < 2147483647LL, SEB_EINVAL, "conv split: T/F/pixel count too large for the packed row index");
  SEB_REQUIRE(s->tc_planes == 2, SEB_EINVAL, "conv split: weight image must have 2 planes");
  dim3 grid((g.M + BM - 1) / BM, s->tc_ntiles);
  conv_split_tc_kernel<NT, STAGES, EK><<<grid, TC_THREADS, SMEM, st>>>(g, reinterpret_cast<const uint8_t*>(s->w_tc));
  SEB_CHECK_LAUNCH("conv_split_tc_kernel");
  return 0;
}

// K-chunk count up to which the persistent kernel is used (SEB200_CONV_PERSIST_MAXK overrides; tuning knob)
int conv_persist_max_chunks() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SEB200_CONV_PERSIST_MAXK"); v = e ? atoi(e) : 12; }   // measured: persistent wins for <= 12 chunks (layers 1-2, conv_2), the 2-CTA ring for longer K
  return v;
}

int conv_y3_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SEB200_CONV_Y3"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

}  // namespace seb

using namespace seb;

extern "C" int seb200_gemm(const SebGemm* s, int engine, void* stream) {
  SEB_REQUIRE(s != nullptr, SEB_EINVAL, "gemm: null descriptor");
  SEB_REQUIRE(s->M > 0 && s->N > 0 && s->K > 0 && s->K % BK == 0, SEB_EINVAL, "gemm: bad sizes M=%d N=%d K=%d", s->M, s->N, s->K);
  SEB_REQUIRE(s->a[0] && s->out && aligned16(s->a[0]) && aligned16(s->out), SEB_EALIGN, "gemm: a/out null or unaligned");
  if (s->loader == SEB_LOAD_ROWS_F16) SEB_REQUIRE(s->lda % 8 == 0 && s->lda >= s->K, SEB_EALIGN, "gemm: fp16 rows need lda %% 8 == 0 and >= K");
  if (s->loader == SEB_LOAD_ROWS || s->loader == SEB_LOAD_ROWS_LN) {
    SEB_REQUIRE(s->lda % 4 == 0 && s->lda >= s->K, SEB_EALIGN, "gemm: lda=%lld must be a multiple of 4 and >= K", s->lda);
  }
  if (s->loader == SEB_LOAD_ROWS2) {
    SEB_REQUIRE(s->K == 128 && s->a[1] && aligned16(s->a[1]) && s->lda % 4 == 0 && s->lda >= 64, SEB_EALIGN, "gemm: the two-source loader needs K == 128, a[1] and lda >= 64");
  }
  if (s->loader == SEB_LOAD_ROWS_LN) {
    SEB_REQUIRE(s->K == 64 && s->ln_gamma && s->ln_beta, SEB_EINVAL, "gemm: LayerNorm loader needs K == 64 and gamma/beta");
  }
  if (s->loader == SEB_LOAD_CONV_ADJ) {
    SEB_REQUIRE((s->lda == 64 || s->lda == 128) && s->nslots == (int)(s->lda / 64) && (s->taps_t == 1 || s->taps_t == 2) && s->stride_f >= 1 && s->dil >= 1,
                SEB_EINVAL, "gemm: bad adjoint-conv geometry");
    SEB_REQUIRE(s->K == s->taps_t * 3 * (int)s->lda, SEB_EINVAL, "gemm: adjoint conv K=%d != taps*lda", s->K);
    SEB_REQUIRE((long long)s->B * s->T * s->Fout == s->M, SEB_EINVAL, "gemm: adjoint conv M != B*T*Fout");
  }
  if (s->loader == SEB_LOAD_CONV || s->loader == SEB_LOAD_CONV_SPLIT) {
    SEB_REQUIRE(s->nslots >= 1 && s->nslots <= 4 && (s->taps_t == 1 || s->taps_t == 2) && s->stride_f >= 1 && s->dil >= 1, SEB_EINVAL, "gemm: bad conv geometry");
    SEB_REQUIRE(s->K == s->taps_t * 3 * s->nslots * 64, SEB_EINVAL, "gemm: conv K=%d != taps*slots*64", s->K);
    SEB_REQUIRE((long long)s->B * s->T * s->Fout == s->M, SEB_EINVAL, "gemm: conv M != B*T*Fout");
    for (int i = 0; i < s->nslots; ++i) SEB_REQUIRE(s->a[i] && aligned16(s->a[i]), SEB_EALIGN, "gemm: conv slot %d null/unaligned", i);
  }
  if (s->loader == SEB_LOAD_HANKEL) {
    SEB_REQUIRE(s->Fin > 0 && s->Fin % 8 == 0 && s->Fin <= s->K && s->stride_f % 4 == 0 && s->lda % 4 == 0 && s->T > 0, SEB_EALIGN, "gemm: bad framing geometry");
  }
  if (s->epilogue == SEB_EPI_GATE) {
    SEB_REQUIRE(s->N % 8 == 0 && s->ldo % 4 == 0, SEB_EALIGN, "gemm: the gate epilogue needs N % 8 == 0 and ldo % 4 == 0");
    SEB_REQUIRE(!s->resid || (s->ldr > 0 && s->ldr <= 2147483647LL && aligned16(s->resid)), SEB_EINVAL, "gemm: gate row bias needs ldr = rows per group > 0");
  }
  if (s->epilogue == SEB_EPI_RESID || s->epilogue == SEB_EPI_RESID_SCALE) SEB_REQUIRE(s->resid && s->ldr % 4 == 0 && aligned16(s->resid), SEB_EALIGN, "gemm: residual null/unaligned");
  if (s->epilogue == SEB_EPI_GLU) SEB_REQUIRE(s->ldo % 2 == 0, SEB_EALIGN, "gemm: ldo must be even");
  else if (s->epilogue == SEB_EPI_GLU_F16) SEB_REQUIRE(s->ldo % 8 == 0 && s->N % 8 == 0, SEB_EALIGN, "gemm: the fp16 GLU epilogue needs ldo %% 8 == 0 and N %% 8 == 0");
  else if (s->epilogue != SEB_EPI_COMPRESS) SEB_REQUIRE(s->ldo % 4 == 0, SEB_EALIGN, "gemm: ldo must be a multiple of 4");
  if (s->epilogue == SEB_EPI_QKV_F16) SEB_REQUIRE(s->N == 192 && s->ldo == 192, SEB_EINVAL, "gemm: the fp16 q|k|v epilogue needs N == ldo == 192");

  const GemmArgs g = to_args(s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int key = s->loader * 16 + s->epilogue;
  if (engine == SEB_ENGINE_SIMT) {
    switch (key) {
      case SEB_LOAD_HANKEL * 16 + SEB_EPI_COMPRESS: return launch_simt<SEB_LOAD_HANKEL, SEB_EPI_COMPRESS>(s, g, st);
      case SEB_LOAD_ROWS * 16 + SEB_EPI_BIAS:       return launch_simt<SEB_LOAD_ROWS, SEB_EPI_BIAS>(s, g, st);
      case SEB_LOAD_HANKEL * 16 + SEB_EPI_BIAS:     return launch_simt<SEB_LOAD_HANKEL, SEB_EPI_BIAS>(s, g, st);
      case SEB_LOAD_ROWS * 16 + SEB_EPI_RESID:      return launch_simt<SEB_LOAD_ROWS, SEB_EPI_RESID>(s, g, st);
      case SEB_LOAD_ROWS * 16 + SEB_EPI_RESID_SCALE: return launch_simt<SEB_LOAD_ROWS, SEB_EPI_RESID_SCALE>(s, g, st);
      case SEB_LOAD_ROWS2 * 16 + SEB_EPI_GATE:      return launch_simt<SEB_LOAD_ROWS2, SEB_EPI_GATE>(s, g, st);
      case SEB_LOAD_CONV * 16 + SEB_EPI_BIAS:       return launch_simt<SEB_LOAD_CONV, SEB_EPI_BIAS>(s, g, st);
      case SEB_LOAD_CONV * 16 + SEB_EPI_SUBPIXEL:   return launch_simt<SEB_LOAD_CONV, SEB_EPI_SUBPIXEL>(s, g, st);
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_SWISH:   return launch_simt<SEB_LOAD_ROWS_LN, SEB_EPI_SWISH>(s, g, st);
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_GLU:     return launch_simt<SEB_LOAD_ROWS_LN, SEB_EPI_GLU>(s, g, st);
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_GLU_F16: return launch_simt<SEB_LOAD_ROWS_LN, SEB_EPI_GLU_F16>(s, g, st);
      case SEB_LOAD_ROWS_F16 * 16 + SEB_EPI_RESID:  return launch_simt<SEB_LOAD_ROWS_F16, SEB_EPI_RESID>(s, g, st);
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_BIAS:    return launch_simt<SEB_LOAD_ROWS_LN, SEB_EPI_BIAS>(s, g, st);
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_QKV_F16: return launch_simt<SEB_LOAD_ROWS_LN, SEB_EPI_QKV_F16>(s, g, st);
      case SEB_LOAD_CONV_ADJ * 16 + SEB_EPI_BIAS:   return launch_simt<SEB_LOAD_CONV_ADJ, SEB_EPI_BIAS>(s, g, st);
      case SEB_LOAD_CONV_ADJ * 16 + SEB_EPI_RESID:  return launch_simt<SEB_LOAD_CONV_ADJ, SEB_EPI_RESID>(s, g, st);
      case SEB_LOAD_CONV * 16 + SEB_EPI_RESID:      return launch_simt<SEB_LOAD_CONV, SEB_EPI_RESID>(s, g, st);
      default: break;
    }
  } else if (engine == SEB_ENGINE_TCGEN05_F32) {
    const int r = launch_train_tc(s, g, st);
    if (r != -100) return r;
  } else if (engine == SEB_ENGINE_TCGEN05) {
    const int nt = s->tc_ntile;
    {
      const int r = launch_tok_gemm(s, g, st);       // persistent kernels for the conformer's token-wise projections
      if (r != -100) return r;
    }
    switch (key) {
      case SEB_LOAD_CONV_SPLIT * 16 + SEB_EPI_BIAS:
        if (nt == 64 && s->tc_ntiles == 1 && s->N == 64 && s->stride_f == 1 && s->Fin == s->Fout && s->taps_t == 2 && s->Fin >= 16 && s->ldo % 4 == 0 && s->tc_planes == 2 && conv_y3_enabled())
          return launch_conv_y3(s, g, st);               // three frequency taps folded into N = 192
        if (nt == 64 && s->tc_ntiles == 1 && s->N == 64 && s->ldo % 4 == 0 && s->tc_planes == 2 && s->Fout >= 64 && s->K / BK <= conv_persist_max_chunks())
          return launch_conv_persist(s, g, st);          // persistent 4-slot ring
        if (nt == 64) return launch_conv_split<64, 2, SEB_EPI_BIAS>(s, g, st);
        break;
      case SEB_LOAD_CONV_SPLIT * 16 + SEB_EPI_SUBPIXEL: if (nt == 128) return launch_conv_split<128, 1, SEB_EPI_SUBPIXEL>(s, g, st); break;
      case SEB_LOAD_HANKEL * 16 + SEB_EPI_COMPRESS: if (nt == 208) return (s->tc_planes == 3) ? launch_tc<208, 1, SEB_LOAD_HANKEL, SEB_EPI_COMPRESS, 8, 1, 3>(s, g, st)
                                                                                             : launch_tc<208, 1, SEB_LOAD_HANKEL, SEB_EPI_COMPRESS>(s, g, st); break;
      case SEB_LOAD_HANKEL * 16 + SEB_EPI_BIAS:     if (nt == 208 && s->tc_planes == 3) return launch_tc<208, 1, SEB_LOAD_HANKEL, SEB_EPI_BIAS, 8, 1, 3>(s, g, st); break;
      case SEB_LOAD_ROWS * 16 + SEB_EPI_BIAS:       if (nt == 208) return (s->tc_planes == 3) ? launch_tc<208, 1, SEB_LOAD_ROWS, SEB_EPI_BIAS, 8, 1, 3>(s, g, st)
                                                                                             : launch_tc<208, 1, SEB_LOAD_ROWS, SEB_EPI_BIAS>(s, g, st); break;
      case SEB_LOAD_ROWS * 16 + SEB_EPI_RESID:      if (nt == 64)  return (s->K <= 128) ? launch_tc<64, 1, SEB_LOAD_ROWS, SEB_EPI_RESID, 4, 4>(s, g, st)
                                                                                      : launch_tc<64, 2, SEB_LOAD_ROWS, SEB_EPI_RESID>(s, g, st); break;
      case SEB_LOAD_ROWS * 16 + SEB_EPI_RESID_SCALE: if (nt == 64) return launch_tc<64, 1, SEB_LOAD_ROWS, SEB_EPI_RESID_SCALE, 4, 4>(s, g, st); break;
      case SEB_LOAD_ROWS2 * 16 + SEB_EPI_GATE:      if (nt == 128) return launch_tc<128, 1, SEB_LOAD_ROWS2, SEB_EPI_GATE>(s, g, st); break;
      case SEB_LOAD_CONV * 16 + SEB_EPI_BIAS:       if (nt == 64)  return launch_tc<64, 2, SEB_LOAD_CONV, SEB_EPI_BIAS>(s, g, st); break;
      case SEB_LOAD_CONV * 16 + SEB_EPI_SUBPIXEL:   if (nt == 128) return launch_tc<128, 1, SEB_LOAD_CONV, SEB_EPI_SUBPIXEL>(s, g, st); break;
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_SWISH:   if (nt == 256) return launch_tc<256, 1, SEB_LOAD_ROWS_LN, SEB_EPI_SWISH>(s, g, st); break;
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_GLU:     if (nt == 256) return launch_tc<256, 1, SEB_LOAD_ROWS_LN, SEB_EPI_GLU>(s, g, st);
                                                    if (nt == 64)  return launch_tc<64, 1, SEB_LOAD_ROWS_LN, SEB_EPI_GLU, 4, 4>(s, g, st); break;
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_GLU_F16: if (nt == 256) return launch_tc<256, 1, SEB_LOAD_ROWS_LN, SEB_EPI_GLU_F16>(s, g, st); break;
      case SEB_LOAD_ROWS_F16 * 16 + SEB_EPI_RESID:  if (nt == 64)  return launch_tc<64, 2, SEB_LOAD_ROWS_F16, SEB_EPI_RESID>(s, g, st); break;
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_BIAS:    if (nt == 192) return launch_tc<192, 1, SEB_LOAD_ROWS_LN, SEB_EPI_BIAS>(s, g, st);
                                                    if (nt == 64)  return launch_tc<64, 1, SEB_LOAD_ROWS_LN, SEB_EPI_BIAS, 4, 4>(s, g, st); break;
      case SEB_LOAD_ROWS_LN * 16 + SEB_EPI_QKV_F16: if (nt == 192) return launch_tc<192, 1, SEB_LOAD_ROWS_LN, SEB_EPI_QKV_F16>(s, g, st);
                                                    if (nt == 64)  return launch_tc<64, 1, SEB_LOAD_ROWS_LN, SEB_EPI_QKV_F16, 4, 4>(s, g, st); break;
      default: break;
    }
  }
  set_error("gemm: loader %d / epilogue %d / engine %d / n-tile %d is not instantiated", s->loader, s->epilogue, engine, s->tc_ntile);
  return SEB_EUNSUPPORTED;
}
