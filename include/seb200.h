/* seb200.h -- C ABI of libseb200.so: the B200 (sm_100a) generator hot path of
 * minyoungpark1/Speech-Enhancement (SCP-GAN / CMGAN TSCNet forward + the
 * power-compressed STFT / iSTFT bracket).
 *
 * The reference has no FFI: its "operator API" for this path is the Python
 * nn.Module protocol (models/generator.py:132-167) plus four DSP helpers
 * (core/function.py:625-703) and predict() (inference_gan.py:75-100).  Each
 * entry point below names the reference code it replaces.  INTEGRATION.md
 * shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: raw device pointers, ints, a cudaStream_t passed as void*.
 *   - every buffer is owned by the caller (PyTorch's caching allocator in the
 *     shipped host code); the library only launches kernels on `stream`, never
 *     synchronises, allocates or frees => CUDA-graph capturable.
 *   - return 0 on success, <0 = SEB_E* argument error, >0 = cudaError_t.
 *     seb200_last_error_string() describes the last failure of this thread.
 *   - activations are fp32, channels-last: [B, T, F, C]; "tokens" means the
 *     flattened [B*T*F, C] view.
 */
#ifndef SEB200_H
#define SEB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SEB200_ABI_VERSION 2

enum {
  SEB_OK = 0,
  SEB_EINVAL = -1,      /* bad shape / null pointer */
  SEB_EALIGN = -2,      /* pointer or stride not 16-byte aligned */
  SEB_EUNSUPPORTED = -3 /* combination not instantiated */
};

/* ---- GEMM engine ------------------------------------------------------------
 * C[M, N] = epilogue( A_loader[M, K] * W[N, K]^T ).  One descriptor covers every
 * dense contraction on the path: DFT / iDFT (core/function.py:690-691,701-702),
 * Conv2d of DilatedDenseNet / conv_2 / SPConvTranspose2d (generator.py:18-20,45,83)
 * as implicit GEMM, and the Linear / pointwise Conv1d layers of ConformerBlock
 * (conformer.py:87-89,137-140,165,169).
 */
enum { SEB_LOAD_ROWS = 0, SEB_LOAD_ROWS_LN = 1, SEB_LOAD_CONV = 2, SEB_LOAD_HANKEL = 3,
       SEB_LOAD_CONV_SPLIT = 4, /* CONV on pre-split inputs: every pixel = 64 bf16 hi + 64 bf16 lo (256 bytes); tcgen05 engine only */
       SEB_LOAD_ROWS2 = 5,      /* K = 128: row m = (a[0][m, 0:64] | a[1][m, 0:64]), both with row stride lda: the two 1x1 convs of
                                   MergeBlock (models/tsc_diffusion.py:21-22,32-34) as one contraction over [x | conditioner] */
       SEB_LOAD_ROWS_F16 = 7,   /* ROWS whose A operand is __half [M, K] (a[0] points at halfs, lda in halfs): the depthwise output v feeding the
                                   pointwise Conv1d(128 -> 64) of the conformer's conv module (conformer.py:169) */
       SEB_LOAD_CONV_ADJ = 6    /* adjoint of CONV (training, dgrad): rows = pixels of the forward conv's input (M = B*T*Fout), a[0] = gradient
                                   image of the forward conv's output [B, T, Fin, lda] (lda = 64 or 128 channels), nslots = lda / 64,
                                   K = taps * lda in (tap, 64-channel part) order; taps_t / dil / stride_f as in the forward conv */ };
enum {
  SEB_EPI_BIAS = 0,     /* out = acc + bias                                  */
  SEB_EPI_SWISH = 1,    /* v = acc + bias; out = v * sigmoid(v)              */
  SEB_EPI_GLU = 2,      /* columns interleaved (value, gate): out[n/2] = v * sigmoid(g) */
  SEB_EPI_RESID = 3,    /* out = alpha * (acc + bias) + resid                */
  SEB_EPI_SUBPIXEL = 4, /* out[(bt*2Fo + 2w + n/64), n%64] = acc + bias      */
  SEB_EPI_COMPRESS = 5, /* columns interleaved (re, im): out[m, k, 0..2] = (|X|^.3, re|X|^-.7, im|X|^-.7) */
  SEB_EPI_QKV_F16 = 6,  /* out is __half [M, 192] = (q * 0.25 * log2(e) | k | v): input of seb200_attention variant 0 */
  SEB_EPI_GATE = 7,     /* columns interleaved (gate, filter): out[n/2] = sigmoid(g) * tanh(f) with (g, f) = acc + bias + rowbias;
                           rowbias = resid[(m / ldr) * N + n] when resid != NULL: one extra bias row per group of ldr consecutive
                           rows (the diffusion-step projection, constant per utterance; models/tsc_diffusion.py:27-37) */
  SEB_EPI_RESID_SCALE = 8, /* out = alpha * (acc + bias + resid): (x + output_residual(y)) / sqrt(2), tsc_diffusion.py:39-41 */
  SEB_EPI_GLU_F16 = 9     /* SEB_EPI_GLU with a __half output [M, N / 2] (ldo in halfs, % 8 == 0): the GLU output u is stored in 16 bits between
                             pw1 -> depthwise (conformer.py:165-166); measured whole-path cost 1.3-2.1e-4 of peak (profiles/r2/fp16_uv_error.txt) */
};
enum { SEB_ENGINE_TCGEN05 = 0, SEB_ENGINE_SIMT = 1,
       SEB_ENGINE_TCGEN05_F32 = 2 /* tcgen05 with THREE bf16 planes per operand (six products, fp32-grade; w_tc packed with planes = 3), fp32
                                     inputs split on the fly: the training step's GEMMs (gradients of this network amplify forward
                                     perturbations ~100x, so the two-plane split of the inference path is not enough there) */ };

typedef struct SebGemm {
  int loader, epilogue;
  int M, N, K;              /* logical sizes; K as packed (multiple of 64)     */
  /* A operand */
  const float* a[4];        /* ROWS*: a[0]; CONV: one pointer per 64-channel slot, newest first; HANKEL: padded signal */
  long long lda;            /* ROWS*: row stride; HANKEL: samples per utterance row of a[0] */
  const float* ln_gamma;    /* ROWS_LN (K == 64): LayerNorm(64) weight / bias, eps 1e-5 */
  const float* ln_beta;
  int B, T, Fin, Fout;      /* CONV geometry: M = B*T*Fout output pixels; HANKEL: M = B*T, hop 100 */
  int taps_t, dil, stride_f, nslots; /* CONV: kernel (taps_t, 3), dilation (dil, 1), stride (1, stride_f), pad (dil*(taps_t-1) top, 1 left/right) */
  /* W operand (see seb200 packing.py / DESIGN.md for the image layouts) */
  const void* w_tc;         /* tcgen05 image: per (n-tile, k-chunk) bf16 planes hi|lo (or hi|mid|lo), 128B-swizzled K-major */
  int tc_ntile;             /* rows per n-tile in w_tc (16..256, multiple of 16) */
  int tc_ntiles;
  int tc_planes;            /* 2: hi+lo, 3 products (network GEMMs); 3: hi+mid+lo, 6 products (DFT / iDFT: fp32-grade) */
  const float* w_simt;      /* fp32 [K][simt_npad] (K-major rows)               */
  int simt_npad;            /* multiple of 64                                   */
  const float* bias;        /* [N] or NULL */
  /* output */
  float* out; long long ldo;
  const float* resid; long long ldr; float alpha;
} SebGemm;

int seb200_gemm(const SebGemm* g, int engine, void* stream);

/* Fused conformer feed-forward half-step on tcgen05 (conformer.py:53-71,128-145,207,210 [+ :211, generator.py:70,72]):
 *   y = x + alpha * (W2 . swish(W1 . LayerNorm(x) + b1) + b2);   out = post_gamma ? LayerNorm_post(y) + resid2 : y
 * x, out, resid2: [tokens, 64] fp32 (out may alias x or resid2); w1_tc / w2_tc: tcgen05 images of W1 [256, 64] and
 * W2 [64, 256] packed with n-tile 64 (see packing.py); the 256-wide hidden activation stays in TMEM / shared memory. */
typedef struct SebFfn {
  const float* x; float* out; long long tokens;
  const float* ln_gamma; const float* ln_beta;
  const void* w1_tc; const float* b1;
  const void* w2_tc; const float* b2;
  float alpha;
  const float* post_gamma; const float* post_beta; const float* resid2;   /* optional (all three or none) */
} SebFfn;
int seb200_ffn_fused(const SebFfn* f, void* stream);

/* ---- DSP bracket --------------------------------------------------------- */
/* predict() glue, inference_gan.py:79-87 + torch.stft's reflect padding: per utterance
 * c = sqrt(L / sum x^2); xpad[b, 0 : Lp+400] = reflect200(wrap_pad(c * x)); c_out[b] = c. */
int seb200_rms_pad(const float* wave, int B, int L, int Lp, int normalize,
                   float* xpad, float* c_out, void* stream);
/* normalize_batch, core/function.py:647-659: the clean utterance takes the NOISY utterance's gain:
 * xpad[b, 0 : Lp+400] = reflect200(wrap_pad(c_in[b] * x)). */
int seb200_scale_pad(const float* wave, int B, int L, int Lp, const float* c_in, float* xpad, void* stream);
/* complex64 (B, F, T) spectrogram (torch.stft layout) -> in3 [B, T, F, 3] = (|x|, re, im); generator.py:146-151 */
int seb200_spec_to_in3(const float* spec_ri, int B, int F, int T, float* in3, void* stream);
/* in3 [B, T, F, 3] -> complex64 (B, F, T): the layout compressed_stft returns (core/function.py:693) */
int seb200_in3_to_spec(const float* in3, int B, int F, int T, float* spec_ri, void* stream);
/* power_uncompress (core/function.py:636-645) of est [B*T, F, 2] into iDFT rows z [B*T, ldz] =
 * (re_0, im_0, re_1, ...), zero padded to ldz */
int seb200_decompress_rows(const float* est, int rows, int F, float* z, int ldz, void* stream);
/* complex64 (B, F, T) -> decompressed iDFT rows z [B*T, ldz] (uncompressed_istft called on its own) */
int seb200_spec_decompress_rows(const float* spec_ri, int B, int F, int T, float* z, int ldz, void* stream);
/* torch.istft overlap-add as a deterministic gather: out[b, m] = sum_t frames[b, t, m + 200 - 100 t] * inv_env[m] * (inv_c ? 1/c[b] : 1) */
int seb200_overlap_add(const float* frames, int B, int T, int ldf, const float* inv_env,
                       const float* c, float* out, int Lout, int ld_out, void* stream);

/* ---- backward of the DSP bracket (consistency-loss chain, core/function.py:231-254) -------------------
 * Gradients of complex tensors follow PyTorch: (dL/dRe, dL/dIm) interleaved.  The DFT contractions of the backward pass are
 * seb200_gemm calls with the transposed bases (ROWS loader for the STFT, HANKEL loader + BIAS epilogue for the iSTFT). */
/* compressed_stft backward, step 1: Y, gY complex64 (B, F, T) -> rows [B*T, ldz] = gX (re, im) per bin, zero padded */
int seb200_compress_backward_rows(const float* spec_ri, const float* gspec_ri, int B, int F, int T, float* rows, int ldz, void* stream);
/* compressed_stft backward, step 3: gframes [B*T, ldf] (after the basis^T GEMM) -> gx [B, L], L = 100*(T-1): the adjoint of
 * torch.stft's reflect padding + framing as a gather */
int seb200_stft_fold(const float* gframes, int B, int T, int ldf, int L, float* gx, void* stream);
/* uncompressed_istft backward, step 1: wpad [B, Lout+400] = zero-pad200(gy * inv_env); its Hankel frames are the frame gradients */
int seb200_istft_grad_pad(const float* gy, const float* inv_env, int B, int Lout, float* wpad, void* stream);
/* uncompressed_istft backward, step 3: gZ rows [B*T, ldz] (after the basis^T GEMM) + Y complex64 (B, F, T) -> gY (B, F, T) */
int seb200_decompress_backward_spec(const float* spec_ri, const float* rows, int B, int F, int T, int ldz, float* gspec_ri, void* stream);

/* ---- encoder / decoder bandwidth kernels ---------------------------------- */
/* DenseEncoder.conv_1[0]: 1x1 conv 3 -> 64 (generator.py:39) on in3 */
int seb200_conv1x1_in3(const float* in3, long long pixels, const float* w /*[64,3]*/,
                       const float* bias, float* out, void* stream);
/* nn.InstanceNorm2d statistics over the (T, F) plane per (b, c): stats[b, c] = (mean, rstd), eps 1e-5, biased var.
 * x: [B, pix_per_b, C]; workspace: doubles, at least seb200_inorm_workspace_bytes() */
long long seb200_inorm_workspace_bytes(int B, long long pix_per_b, int C);
int seb200_inorm_stats(const float* x, int B, long long pix_per_b, int C, float* stats,
                       void* workspace, long long workspace_bytes, void* stream);
/* y = PReLU(gamma * (x - mean) * rstd + beta): InstanceNorm2d(affine) + PReLU(C) (generator.py:21-22,40-41).
 * out_format 0: y is fp32 [B, pix, 64]; 1: y is the pre-split conv-input format (per pixel 64 bf16 hi | 64 bf16 lo) */
int seb200_inorm_prelu(const float* x, int B, long long pix_per_b, int C, const float* stats,
                       const float* gamma, const float* beta, const float* slope, void* y, int out_format, void* stream);
/* fp32 [pixels, 64] -> pre-split conv-input format (the decoders read the TSCB output this way) */
int seb200_split_planes(const float* x, long long pixels, void* y, void* stream);
/* MaskDecoder.conv_1: Conv2d(64 -> 1, (1,2)) on [B*T, 202, 64] -> raw [B*T, 201] (generator.py:100) */
int seb200_mask_conv(const float* x, long long rows, int Fin, const float* w /*[2][64]: tap-major*/,
                     float bias, float* out, void* stream);
/* ComplexDecoder: InstanceNorm(64)+PReLU(64) applied on load, then Conv2d(64 -> 2, (1,2)) (generator.py:127-128) */
int seb200_complex_conv(const float* x, int B, long long rows_per_b, int Fin, const float* stats,
                        const float* gamma, const float* beta, const float* slope,
                        const float* w /*[2 out][2 taps][64]*/, const float* bias /*[2]*/, float* out /*[rows, 201, 2]*/, void* stream);
/* Mask tail + recombination (generator.py:110-112,158-165): mask = PReLU_f(wf * PReLU(IN(raw)) + bf);
 * est[.., 0..1] = mask * in3[.., 1..2] + cplx */
int seb200_mask_recombine(const float* mask_raw, const float* mask_stats /*[B,1,2]*/, int B, long long rows_per_b, int F,
                          float in_gamma, float in_beta, float slope1, float wf, float bf, const float* slope_f /*[F]*/,
                          const float* in3, const float* cplx, float* est /*[rows, F, 2]*/, float* mask_out /*optional [rows,F]*/, void* stream);
/* est [B*T, F, 2] -> two fp32 (B, 1, T, F) tensors (identical memory order: a de-interleave) */
int seb200_split_ri(const float* est, long long n, float* re, float* im, void* stream);

/* ---- conformer kernels ---------------------------------------------------- */
/* Token addressing of a sequence set: token(seq, i) = (seq / inner) * outer_stride + (seq % inner) + i * pos_stride.
 * time conformer (generator.py:69): inner = F', outer_stride = T*F', pos_stride = F', nseq = B*F', n = T
 * freq conformer (generator.py:71): inner = 1,  outer_stride = F',   pos_stride = 1,  nseq = B*T,  n = F' */
typedef struct SebSeq { int nseq, n, inner; long long outer_stride, pos_stride; } SebSeq;

/* Attention core with Shaw relative positions (conformer.py:103-122): qkv [tokens, 192] = (q | k | v), heads 4 x 16;
 * out [tokens, 64] fp32 ('b h n d -> b n (h d)').
 *   variant 0 (tensor cores): qkv is __half, q pre-scaled by 0.25 * log2(e) (SEB_EPI_QKV_F16 writes it);
 *                             rel_pos_emb_h = rel_pos_emb [1025, 16] rounded to IEEE fp16 with the 16 halfs of every row
 *                             in MMA-fragment order k = 0,1,8,9, 2,3,10,11, 4,5,12,13, 6,7,14,15 (packed once by the host)
 *   variant 3 (tcgen05): same inputs as variant 0.  S = Q K^T (fp32) and R = Q E_window^T (fp16) accumulate in tensor memory,
 *                        the per-row skew of R goes through thread-private shared-memory rows, P is the TMEM A operand of
 *                        the P V product; key tiles beyond the +-512 clamp skip the rel-pos GEMM (attention_tc.cu)
 *   variant 1 (fp32 SIMT cross-check): qkv is float, unscaled; rel_pos_emb [1025, 16] fp32 */
int seb200_attention(const void* qkv, const float* rel_pos_emb, const void* rel_pos_emb_h, const SebSeq* seq, float* out,
                     int variant, void* stream);
/* DepthWiseConv1d(128, k=31, pad 15/15) + BatchNorm1d(eval) + Swish along the sequence axis (conformer.py:166-168):
 * x, y [tokens, 128]; w [31][128] (tap-major); bn_scale/bn_shift fold conv bias, running stats and affine */
int seb200_dwconv_bn_swish(const float* x, const SebSeq* seq, const float* w, const float* bn_scale,
                           const float* bn_shift, float* y, void* stream);
/* the same stage with y stored as __half [tokens, 128] (fp32 arithmetic inside); x is float, or __half when x_is_half != 0 */
int seb200_dwconv_bn_swish_f16(const void* x, int x_is_half, const SebSeq* seq, const float* w, const float* bn_scale,
                               const float* bn_shift, void* y, void* stream);
/* post_norm + the TSCB outer residual (conformer.py:211, generator.py:70,72): out = LN(x) * g + b + resid */
int seb200_layernorm_residual(const float* x, long long tokens, const float* gamma, const float* beta,
                              const float* resid, float* out, void* stream);

/* ---- training step (SURVEY 8f row f1 / 8e training): train-mode forward pieces and the backward of every op of the generator ----------
 * Reference: model(noisy_spec) under model.train() (core/function.py:221), loss.backward() (:274), SyncBatchNorm (main_gan.py:154-155),
 * DDP gradient all-reduce (main_gan.py:168-171).  Dense contractions go through seb200_gemm (forward: the layer's own descriptor; dgrad:
 * the same contraction with the transposed weight image, or SEB_LOAD_CONV_ADJ) with engine SEB_ENGINE_TCGEN05_F32 or SEB_ENGINE_SIMT;
 * weight gradients through seb200_wgrad.  Weight images are rebuilt on the device from the live parameters every step
 * (seb200_pack_weights_device).  Reductions are two-stage and deterministic (no atomics on global memory); `workspace` buffers are scratch
 * of at least seb200_train_workspace_floats() floats (or doubles where the parameter is double*), unless a size query is named. */
long long seb200_train_workspace_floats(void);

/* W[n, k] = w[n * sn + (k / n1) * s0 + (k % n1) * s1] (k < K) packed into the tcgen05 image (planes 2 or 3) and / or the K-major fp32
 * image, all in DEVICE memory; same bytes as seb200_pack_weights.  The index map covers Linear weights and their transposes (dgrad) and
 * Conv2d weights [Cout, Cin, kt, kf] in the engine's K order (tap, cin) and their per-slot adjoints (csrc/pack_dev.cu). */
int seb200_pack_weights_device(const float* w, int N, int K, int n1, long long sn, long long s0, long long s1, int tc_ntile, int planes,
                               void* w_tc, float* w_simt, void* stream);
/* The same for njobs images in ceil(njobs / 24) launches: `jobs` is a HOST array of descriptors holding the arguments above (device pointers).
 * A training step re-packs 176 images after every optimizer step (core/function.py:277 changes the parameters in place). */
typedef struct SebPackJob {
  const float* w; void* w_tc; float* w_simt;
  long long sn, s0, s1;
  int N, K, n1, tc_ntile, planes, reserved;
} SebPackJob;
int seb200_pack_weights_device_batch(const SebPackJob* jobs, int njobs, void* stream);

/* dW[n, k] = sum_m g_out[m, n] * A[m, k], db[n] = sum_m g_out[m, n]: `a` is the FORWARD GEMM's descriptor (loader ROWS / ROWS_LN / CONV, a[],
 * lda, ln_*, conv geometry, M, K; weight / output fields ignored), g_out [M, N] with row stride ldg, N a multiple of 64 (<= 256).
 * dW lands at dw[n * sn + (k / n1) * s0 + (k % n1) * s1] for k < k_logical (the parameter's own layout); db [N] or NULL. */
int seb200_wgrad_splits(int M, int N, int K);
long long seb200_wgrad_workspace_floats(int M, int N, int K);
int seb200_wgrad(const SebGemm* a, const float* g_out, long long ldg, int N, int k_logical, int n1, long long sn, long long s0, long long s1,
                 float* dw, float* db, float* workspace, long long workspace_floats, void* stream);

/* nn.Dropout keep-mask (conformer.py:125,139,141), Philox4x32-10: mask[i] = (u_i >= p), u_i from counter offset + i / 4, key seed.  Every
 * consumer below takes the mask as an explicit pointer (uint8, 1 = keep; NULL = no dropout), so a caller may inject its own draw. */
int seb200_dropout_mask(unsigned char* mask, long long n, float p, unsigned long long seed, unsigned long long offset, void* stream);
/* h = swish(a) * keep * scale (conformer.py:138-139); backward da = dh * keep * scale * swish'(a).  n % 4 == 0. */
int seb200_swish_dropout(const float* a, const unsigned char* mask, float scale, float* h, long long n, void* stream);
int seb200_swish_dropout_bwd(const float* a, const unsigned char* mask, float scale, const float* dh, float* da, long long n, void* stream);
/* y = resid + scale * keep * t (dropout, Scale(0.5) and the block residual, conformer.py:60,141,207-210; resid may alias y);
 * backward of the branch: dt = scale * keep * dy */
int seb200_dropout_residual(const float* t, const unsigned char* mask, float scale, const float* resid, float* y, long long n, void* stream);
int seb200_scale_mask(const float* dy, const unsigned char* mask, float scale, float* dt, long long n, void* stream);
/* GLU in the natural channel order (conformer.py:30-38): a [M, 2C] = (value | gate) -> u [M, C] = value * sigmoid(gate); backward da [M, 2C] */
int seb200_glu(const float* a, long long M, int C, float* u, void* stream);
int seb200_glu_bwd(const float* a, const float* du, long long M, int C, float* da, void* stream);
/* LayerNorm(64) backward (conformer.py:67,162,204): dx = (add ? add : 0) + dLN(x; dy) (add may alias dx), dgamma / dbeta [64] */
int seb200_layernorm_bwd(const float* x, const float* gamma, const float* dy, const float* add, float* dx, long long tokens,
                         float* dgamma, float* dbeta, float* workspace, void* stream);
/* BatchNorm1d(128) in train mode (conformer.py:167) on c [M, 128].  seb200_bn_sums: LOCAL sums[0:128] = sum c, [128:256] = sum c^2 (fp64).
 * Under SyncBatchNorm (main_gan.py:154) the caller all-reduces `sums` and the token count across ranks, then seb200_bn_finalize turns
 * them into batch mean | rstd (mean_rstd [256]), the folded scale | shift (scale_shift [256]) and updates running_mean / running_var
 * (momentum, unbiased variance) and num_batches_tracked (any of the three may be NULL).  seb200_bn_swish: v = swish(scale * c + shift). */
int seb200_bn_sums(const float* c, long long M, double* sums, double* workspace, void* stream);
int seb200_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float momentum, float eps, float* scale_shift, float* mean_rstd, void* stream);
int seb200_bn_swish(const float* c, long long M, const float* scale_shift, float* v, void* stream);
/* backward of v = swish(BN(c)): LOCAL sums[0:128] = sum dz, [128:256] = sum dz * chat (all-reduced under SyncBatchNorm), then
 * dc = gamma * rstd * (dz - S1 / count - chat * S2 / count) with the global sums; dgamma = S2, dbeta = S1 of `sums_local` (this rank's own
 * sums as SyncBatchNorm returns them -- the data-parallel gradient all-reduce adds the ranks; NULL = sums) */
int seb200_bn_swish_bwd_sums(const float* c, const float* dv, long long M, const float* scale_shift, const float* mean_rstd, double* sums,
                             double* workspace, void* stream);
int seb200_bn_swish_bwd_apply(const float* c, const float* dv, long long M, const float* scale_shift, const float* mean_rstd, const double* sums,
                              const double* sums_local, double count, float* dc, float* dgamma, float* dbeta, void* stream);
/* y = scale * DWConv31(x; w) + shift per channel, no activation (w [31][128] tap-major): the train-mode forward of DepthWiseConv1d
 * (conformer.py:40-48) with scale = 1, shift = bias, and its data gradient with reversed taps, scale = 1, shift = 0 */
int seb200_dwconv(const float* x, const SebSeq* seq, const float* w, const float* scale, const float* shift, float* y, void* stream);
/* depthwise weight gradient in the parameter layout (128, 1, 31) and bias gradient [128] from the input u and the output gradient dc */
int seb200_dwconv_wgrad(const float* u, const float* dc, const SebSeq* seq, float* dw, float* db, float* workspace, void* stream);
/* InstanceNorm2d(affine) + PReLU backward (generator.py:21-22,40-41,46-47,101-102,120-121), C = 64 or 1: x the normalisation's input,
 * stats from seb200_inorm_stats, dy the gradient of the PReLU output -> dx, dgamma / dbeta / dslope [C] */
long long seb200_inorm_bwd_workspace_doubles(int B, long long pix_per_b, int C);
int seb200_inorm_prelu_bwd(const float* x, const float* dy, int B, long long pix_per_b, int C, const float* stats, const float* gamma,
                           const float* beta, const float* slope, float* dx, float* dgamma, float* dbeta, float* dslope,
                           double* workspace, long long workspace_doubles, void* stream);
/* decoder heads Conv2d(64 -> NO, (1, 2)), NO = 1 (MaskDecoder.conv_1) or 2 (ComplexDecoder.conv), generator.py:100,122, with the weight
 * in the PARAMETER layout (NO, 64, 1, 2) and the bias by pointer: x [rows, Fin, 64] -> out [rows, Fin - 1, NO]; backward dx, dw, db */
int seb200_head_conv(const float* x, long long rows, int Fin, const float* w, const float* bias, int NO, float* out, void* stream);
int seb200_head_conv_bwd(const float* x, const float* dout, long long rows, int Fin, const float* w, int NO, float* dx, float* dw, float* db,
                         float* workspace, void* stream);
/* seb200_mask_recombine with the five scalars read through device pointers (norm.weight, norm.bias, prelu.weight, final_conv.weight,
 * final_conv.bias: they change every optimizer step), and the backward of  est = PReLU_f(wf * p1 + bf) * (re, im) + cplx  w.r.t.
 * p1 = PReLU(IN(mask_raw)): dp1 [B*T, F], dslope_f [F], dwf, dbf (the gradient w.r.t. cplx is d_est itself) */
int seb200_mask_recombine_dev(const float* mask_raw, const float* mask_stats, int B, long long rows_per_b, int F, const float* const* scalars5,
                              const float* slope_f, const float* in3, const float* cplx, float* est, void* stream);
int seb200_mask_tail_bwd(const float* mask_raw, const float* mask_stats, int B, long long rows_per_b, int F, const float* const* scalars5,
                         const float* slope_f, const float* in3, const float* dest, float* dp1, float* dslope_f, float* dwf, float* dbf,
                         float* workspace, void* stream);
/* DenseEncoder.conv_1[0] weight gradient: dw [64][3], db [64] from in3 [pixels, 3] and the output gradient g [pixels, 64] */
int seb200_conv1x1_in3_wgrad(const float* in3, const float* g, long long pixels, float* dw, float* db, float* workspace, void* stream);
/* (d final_real, d final_imag) (B, 1, T, F) -> d_est [B*T, F, 2]: the inverse of seb200_split_ri */
int seb200_merge_ri(const float* re, const float* im, long long n, float* est, void* stream);
/* fp32 q|k|v [M, 192] -> the fp16 copy (q scaled by 0.25 * log2 e) seb200_attention variants 0 / 3 read */
int seb200_qkv_to_f16(const float* qkv, long long M, void* out, void* stream);
/* fp32-grade attention core for training (3xTF32 mma.sync): qkv fp32 [tokens, 192] unscaled -> out [tokens, 64] and lse [tokens, 4];
 * backward from (qkv, out, lse, dout) to dqkv [tokens, 192] and the gradient of rel_pos_emb [1025, 16] under the +-512 clamp */
int seb200_attention_train_fwd(const float* qkv, const float* rel_pos_emb, const SebSeq* seq, float* out, float* lse, void* stream);
long long seb200_attention_bwd_workspace_floats(long long tokens);
int seb200_attention_bwd(const float* qkv, const float* rel_pos_emb, const SebSeq* seq, long long tokens, const float* out, const float* lse,
                         const float* dout, float* dqkv, float* drel, float* workspace, long long workspace_floats, void* stream);

/* ---- diffusion variant (models/tsc_diffusion.py) ------------ */
/* MergeBlock's diffusion-step branch (tsc_diffusion.py:27-29 + models/DiffuSE.py:46-62) folded into a per-step bias of the
 * merge GEMM:  e = table[step] (integer step) or lerp(table[floor], table[ceil]) (fractional step);
 *   h = silu(W1 e + b1); h = silu(W2 h + b2); d = Wp h + bp   (128 -> 512 -> 512 -> 64)
 *   d_out[s, 0:64] = d;  rowbias[s, n] = sum_k wm[n, k] * d[k]  (n < 128; wm = merge_diffusion weight in the row order the
 *   merge GEMM uses), so that W_m (x + d) = W_m x + rowbias.  nsteps entries (1 or B), max_steps rows in table [max_steps, 128].
 * All device pointers; fp32. */
int seb200_diffusion_embed(const float* steps, int nsteps, const float* table, int max_steps,
                           const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* wp, const float* bp, const float* wm,
                           float* d_out, float* rowbias, void* stream);

/* One waveform update of the reverse process, predict_tsc in inference_diffuse.py:255-264:
 *   out[b, i] = (ca * audio[b, i] + cb * noisy[b, i] + cc * pred[b, i] + cs * noise[b, i]) * (c_div ? 1 / c_div[b] : 1)
 * audio / pred / noise / out: [B, L] contiguous (noise may be NULL); noisy: rows of stride ld_noisy (a view of the padded
 * conditioning buffer); c_div: the per-utterance RMS gain removed after the last step (:265), or NULL. */
int seb200_diffusion_update(const float* audio, const float* noisy, long long ld_noisy, const float* pred, const float* noise,
                            int B, long long L, float ca, float cb, float cc, float cs, const float* c_div, float* out, void* stream);

/* ---- host-side helpers (no launches) ------------------------------------------ */
/* Sizes of the two weight images of one logical W [N, K] (see SebGemm.w_tc / w_simt): bytes of the tcgen05 image, floats of the
 * fp32 image, the padded K, the number of n-tiles and the padded N of the fp32 image.  Any output pointer may be NULL. */
int seb200_packed_weight_sizes(int N, int K, int tc_ntile, int planes, long long* tc_bytes, long long* simt_floats,
                               int* k_padded, int* tc_ntiles, int* simt_npad);
/* Pack W [N, K] fp32 (HOST memory, row-major: nn.Linear / reshaped conv weight) into the tcgen05 image (bf16 hi | lo [| mid] planes,
 * 128-byte-swizzled K-major blocks per (n-tile, 64-wide k-chunk)) and / or the K-major fp32 image; both in HOST memory, sized by
 * seb200_packed_weight_sizes; the caller copies them to the device.  planes: 2 = network GEMMs, 3 = DFT / iDFT bases. */
int seb200_pack_weights(const float* w, int N, int K, int tc_ntile, int planes, void* w_tc, float* w_simt);
/* Bytes of activation workspace one forward over B utterances x T frames x F bins needs, as the shipped host allocates it
 * (kind 0: generator, models/generator.py; kind 1: diffusion variant, models/tsc_diffusion.py); -1 on bad arguments. */
long long seb200_workspace_bytes(int kind, int B, int T, int F);

/* ---- misc ------------------------------------------------------------------ */
int seb200_version(void);
const char* seb200_last_error_string(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long seb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SEB200_H */
