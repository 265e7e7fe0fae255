"""Tensor-level wrappers over the C ABI (one function per exported kernel).

These are the only places where torch tensors meet ``libseb200.so``: every
wrapper checks device/dtype/contiguity, passes raw pointers plus the current
CUDA stream, and turns a non-zero return code into ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TCGEN05_F32, EPI_BIAS, EPI_COMPRESS, EPI_GLU, EPI_QKV_F16, EPI_RESID, EPI_SUBPIXEL, EPI_SWISH,
                   LOAD_CONV, LOAD_CONV_SPLIT, LOAD_HANKEL, LOAD_ROWS, LOAD_ROWS_LN, SebFfn, SebGemm, SebSeq, check, ptr, require_cuda, stream_ptr)
from .packing import PackedWeight

ENGINES = {"tcgen05": ENGINE_TCGEN05, "simt": ENGINE_SIMT, "tcgen05_f32": ENGINE_TCGEN05_F32}


class Profiler:
    """Per-launch CUDA-event timing on the launching stream, keyed by a kernel label.

    ``with ops.profile(only={"dconv"}) as prof: ...`` brackets every matching launch with two events
    (a few microseconds of host work each); ``prof.summary()`` returns per-label launch counts, total
    milliseconds and the algorithmic FLOPs / bytes the wrappers attached to each launch."""

    def __init__(self, only=None):
        self.only = set(only) if only else None
        self.records = []

    def begin(self, label, flops, nbytes):
        if self.only is not None and label not in self.only:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (label, flops, nbytes, e0)

    def end(self, tok):
        if tok is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append(tok + (e1,))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for label, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(label, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out

    def __enter__(self):
        global _PROF
        self._prev, _PROF = _PROF, self
        return self

    def __exit__(self, *exc):
        global _PROF
        _PROF = self._prev
        return False


_PROF = None


def profile(only=None) -> Profiler:
    return Profiler(only)


def _pb(label, flops=0.0, nbytes=0.0):
    return _PROF.begin(label, flops, nbytes) if _PROF is not None else None


def _pe(tok):
    if tok is not None:
        _PROF.end(tok)


def default_engine() -> str:
    return os.environ.get("SEB200_ENGINE", "tcgen05")


def _f32c(*ts):
    for t in ts:
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"expected contiguous float32 tensors, got {t.dtype} contiguous={t.is_contiguous()}")
    require_cuda(*ts)


def gemm(*, loader: int, epilogue: int, M: int, w: PackedWeight, a: Sequence[torch.Tensor], out: torch.Tensor,
         ldo: int, N: Optional[int] = None, lda: int = 0, ln: Optional[tuple] = None, conv: Optional[dict] = None,
         resid: Optional[torch.Tensor] = None, ldr: int = 0, alpha: float = 1.0, engine: str = "tcgen05",
         label: str = "gemm", k_logical: Optional[int] = None):
    """One launch of the GEMM engine (see include/seb200.h: SebGemm)."""
    lib = _lib.load()
    if epilogue in (EPI_QKV_F16, _lib.EPI_GLU_F16):
        if out.dtype != torch.float16 or not out.is_contiguous():
            raise RuntimeError("the fp16 epilogues (q|k|v, GLU) write a contiguous float16 tensor")
        _f32c(resid, *a)
    elif loader == _lib.LOAD_ROWS_F16:
        if a[0].dtype != torch.float16 or not a[0].is_contiguous() or not a[0].is_cuda:
            raise RuntimeError("LOAD_ROWS_F16 reads a contiguous CUDA float16 [M, K] tensor")
        _f32c(out, resid)
    elif loader == LOAD_CONV_SPLIT:
        for t in a:
            if t.dtype != torch.bfloat16 or not t.is_contiguous() or not t.is_cuda:
                raise RuntimeError("LOAD_CONV_SPLIT reads contiguous CUDA bfloat16 [pixels, 2, 64] (hi | lo) tensors")
        _f32c(out, resid)
    else:
        _f32c(out, resid, *a)
    g = SebGemm()
    g.loader, g.epilogue = loader, epilogue
    g.M, g.N, g.K = M, (w.N if N is None else N), w.K
    for i in range(4):
        g.a[i] = ptr(a[i]) if i < len(a) else 0
    g.lda = lda
    if ln is not None:
        g.ln_gamma, g.ln_beta = ptr(ln[0]), ptr(ln[1])
    if conv is not None:
        g.B, g.T, g.Fin, g.Fout = conv["B"], conv["T"], conv["Fin"], conv["Fout"]
        g.taps_t, g.dil, g.stride_f, g.nslots = conv.get("taps_t", 1), conv.get("dil", 1), conv.get("stride_f", 1), conv.get("nslots", 1)
    g.w_tc, g.tc_ntile, g.tc_ntiles, g.tc_planes = ptr(w.w_tc), w.tc_ntile, w.tc_ntiles, w.planes
    g.w_simt, g.simt_npad = ptr(w.w_simt), w.simt_npad
    g.bias = ptr(w.bias)
    g.out, g.ldo = ptr(out), ldo
    g.resid, g.ldr, g.alpha = ptr(resid), ldr, alpha
    if _PROF is not None:
        # algorithmic bytes: the A rows once, the output once (fp16 q|k|v: 2 B, GLU / gate: half the columns), the residual once
        n_out = g.N // 2 if epilogue in (EPI_GLU, _lib.EPI_GATE, _lib.EPI_GLU_F16) else g.N
        nbytes = (2.0 if loader == _lib.LOAD_ROWS_F16 else 4.0) * M * (k_logical or w.K) + (2.0 if epilogue in (EPI_QKV_F16, _lib.EPI_GLU_F16) else 4.0) * M * n_out
        if epilogue in (EPI_RESID, _lib.EPI_RESID_SCALE):
            nbytes += 4.0 * M * g.N
        tok = _pb(label, 2.0 * M * g.N * (k_logical or w.K), nbytes)
    else:
        tok = None
    check(lib.seb200_gemm(C.byref(g), ENGINES[engine], stream_ptr()), "seb200_gemm")
    _pe(tok)
    return out


def diffusion_embed(steps: torch.Tensor, table, w1, b1, w2, b2, wp, bp, wm, d_out: torch.Tensor, rowbias: torch.Tensor):
    """MergeBlock's diffusion-step branch (tsc_diffusion.py:27-29, DiffuSE.py:46-62): steps float32 [n] -> d_out [n, 64] and
    rowbias [n, 128] = wm . d (the per-utterance bias row of the merge GEMM)."""
    _f32c(steps, table, w1, b1, w2, b2, wp, bp, wm, d_out, rowbias)
    n = steps.numel()
    if table.shape[1] != 128 or tuple(d_out.shape) != (n, 64) or tuple(rowbias.shape) != (n, 128):
        raise RuntimeError("diffusion_embed: bad shapes")
    tok = _pb("diffusion_embed")
    check(_lib.load().seb200_diffusion_embed(ptr(steps), n, ptr(table), table.shape[0], ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(wp), ptr(bp),
                                             ptr(wm), ptr(d_out), ptr(rowbias), stream_ptr()), "seb200_diffusion_embed")
    _pe(tok)
    return d_out, rowbias


# ---- backward of the DSP bracket ----------------------------------------------------------------
def compress_backward_rows(spec: torch.Tensor, gspec: torch.Tensor, rows: torch.Tensor):
    """Y, gY complex64 (B, F, T) -> rows [B*T, ldz] = gX (re, im), zero padded."""
    require_cuda(spec, gspec)
    _f32c(rows)
    if spec.dtype != torch.complex64 or gspec.dtype != torch.complex64 or spec.shape != gspec.shape:
        raise RuntimeError("compress_backward_rows: spec / gspec must be complex64 of the same shape")
    B, F, T = spec.shape
    tok = _pb("compress_bwd", 0.0, 24.0 * B * F * T) if _PROF is not None else None
    check(_lib.load().seb200_compress_backward_rows(ptr(torch.view_as_real(spec.contiguous())), ptr(torch.view_as_real(gspec.contiguous())), B, F, T,
                                                    ptr(rows), rows.shape[1], stream_ptr()), "seb200_compress_backward_rows")
    _pe(tok)
    return rows


def stft_fold(gframes: torch.Tensor, B: int, T: int, L: int):
    _f32c(gframes)
    gx = torch.empty(B, L, device=gframes.device, dtype=torch.float32)
    tok = _pb("stft_fold", 0.0, 4.0 * gframes.numel()) if _PROF is not None else None
    check(_lib.load().seb200_stft_fold(ptr(gframes), B, T, gframes.shape[1], L, ptr(gx), stream_ptr()), "seb200_stft_fold")
    _pe(tok)
    return gx


def istft_grad_pad(gy: torch.Tensor, inv_env: torch.Tensor):
    _f32c(gy, inv_env)
    B, Lout = gy.shape
    wpad = torch.empty(B, Lout + 400, device=gy.device, dtype=torch.float32)
    tok = _pb("istft_grad_pad", 0.0, 8.0 * gy.numel()) if _PROF is not None else None
    check(_lib.load().seb200_istft_grad_pad(ptr(gy), ptr(inv_env), B, Lout, ptr(wpad), stream_ptr()), "seb200_istft_grad_pad")
    _pe(tok)
    return wpad


def decompress_backward_spec(spec: torch.Tensor, rows: torch.Tensor):
    """Y complex64 (B, F, T) + gZ rows [B*T, ldz] -> gY complex64 (B, F, T)."""
    require_cuda(spec)
    _f32c(rows)
    if spec.dtype != torch.complex64:
        raise RuntimeError("decompress_backward_spec: spec must be complex64")
    B, F, T = spec.shape
    g = torch.empty(B, F, T, device=spec.device, dtype=torch.complex64)
    tok = _pb("decompress_bwd", 0.0, 24.0 * B * F * T) if _PROF is not None else None
    check(_lib.load().seb200_decompress_backward_spec(ptr(torch.view_as_real(spec.contiguous())), ptr(rows), B, F, T, rows.shape[1],
                                                      ptr(torch.view_as_real(g)), stream_ptr()), "seb200_decompress_backward_spec")
    _pe(tok)
    return g


def diffusion_update(audio, noisy, pred, noise, ca: float, cb: float, cc: float, cs: float, c_div=None, out=None):
    """out = (ca * audio + cb * noisy + cc * pred + cs * noise) [/ c_div per utterance]; noisy may be a row-strided view."""
    _f32c(audio, pred, noise, c_div)
    require_cuda(noisy)
    B, L = audio.shape
    if noisy.dtype != torch.float32 or noisy.shape != audio.shape or noisy.stride(1) != 1 or pred.shape != audio.shape:
        raise RuntimeError("diffusion_update: noisy / pred must be float32 (B, L) like audio (noisy may have a row stride)")
    if out is None:
        out = torch.empty_like(audio)
    tok = _pb("diffusion_update", 0.0, 20.0 * audio.numel()) if _PROF is not None else None
    check(_lib.load().seb200_diffusion_update(ptr(audio), ptr(noisy), noisy.stride(0), ptr(pred), ptr(noise), B, L, ca, cb, cc, cs, ptr(c_div),
                                              ptr(out), stream_ptr()), "seb200_diffusion_update")
    _pe(tok)
    return out


def ffn_fused(x, out, ln, w1: PackedWeight, w2: PackedWeight, alpha: float = 0.5, post=None, resid2=None):
    """y = x + alpha * FF(LN(x)); with ``post=(gamma, beta)``: out = LN_post(y) + resid2.  One tcgen05 kernel."""
    _f32c(x, out, resid2)
    if w1.tc_ntile != 64 or w2.tc_ntile != 64 or w1.N != 256 or w2.N != 64 or w1.K != 64 or w2.K != 256:
        raise RuntimeError("ffn_fused expects W1 [256,64] and W2 [64,256] packed with n-tile 64")
    f = SebFfn()
    f.x, f.out, f.tokens = ptr(x), ptr(out), x.numel() // 64
    f.ln_gamma, f.ln_beta = ptr(ln[0]), ptr(ln[1])
    f.w1_tc, f.b1, f.w2_tc, f.b2 = ptr(w1.w_tc), ptr(w1.bias), ptr(w2.w_tc), ptr(w2.bias)
    f.alpha = alpha
    if post is not None:
        f.post_gamma, f.post_beta, f.resid2 = ptr(post[0]), ptr(post[1]), ptr(resid2)
    tok = _pb("ffn_fused", 2.0 * 2 * 64 * 256 * f.tokens, 4.0 * 64 * f.tokens * (3 if post is None else 4)) if _PROF is not None else None
    check(_lib.load().seb200_ffn_fused(C.byref(f), stream_ptr()), "seb200_ffn_fused")
    _pe(tok)
    return out


def rms_pad(wave: torch.Tensor, Lp: int, normalize: bool = True):
    _f32c(wave)
    B, L = wave.shape
    xpad = torch.empty(B, Lp + 400, device=wave.device, dtype=torch.float32)
    c = torch.empty(B, device=wave.device, dtype=torch.float32)
    tok = _pb("rms_pad", 0.0, 8.0 * wave.numel()) if _PROF is not None else None
    check(_lib.load().seb200_rms_pad(ptr(wave), B, L, Lp, int(normalize), ptr(xpad), ptr(c), stream_ptr()), "seb200_rms_pad")
    _pe(tok)
    return xpad, c


def scale_pad(wave: torch.Tensor, Lp: int, c: torch.Tensor):
    """xpad[b] = reflect200(wrap_pad(c[b] * wave[b])) with the caller's per-utterance gain (normalize_batch's clean branch)."""
    _f32c(wave); _f32c(c)
    B, L = wave.shape
    if c.numel() != B:
        raise RuntimeError("scale_pad: one gain per utterance expected")
    xpad = torch.empty(B, Lp + 400, device=wave.device, dtype=torch.float32)
    tok = _pb("rms_pad", 0.0, 8.0 * wave.numel()) if _PROF is not None else None
    check(_lib.load().seb200_scale_pad(ptr(wave), B, L, Lp, ptr(c), ptr(xpad), stream_ptr()), "seb200_scale_pad")
    _pe(tok)
    return xpad


def spec_to_in3(spec: torch.Tensor, out: Optional[torch.Tensor] = None):
    """complex64 (B, F, T) -> in3 [B, T, F, 3]."""
    require_cuda(spec)
    if spec.dtype != torch.complex64:
        raise RuntimeError("spectrogram must be complex64")
    sr = torch.view_as_real(spec.contiguous())
    B, F, T = spec.shape
    if out is None:
        out = torch.empty(B, T, F, 3, device=spec.device, dtype=torch.float32)
    tok = _pb("spec_to_in3", 0.0, 20.0 * B * F * T) if _PROF is not None else None
    check(_lib.load().seb200_spec_to_in3(ptr(sr), B, F, T, ptr(out), stream_ptr()), "seb200_spec_to_in3")
    _pe(tok)
    return out


def in3_to_spec(in3: torch.Tensor):
    _f32c(in3)
    B, T, F, _ = in3.shape
    out = torch.empty(B, F, T, 2, device=in3.device, dtype=torch.float32)
    tok = _pb("in3_to_spec", 0.0, 20.0 * B * F * T) if _PROF is not None else None
    check(_lib.load().seb200_in3_to_spec(ptr(in3), B, F, T, ptr(out), stream_ptr()), "seb200_in3_to_spec")
    _pe(tok)
    return torch.view_as_complex(out)


def decompress_rows(est: torch.Tensor, z: torch.Tensor):
    _f32c(est, z)
    F = est.shape[-2]                      # est: [..., F, 2]
    rows = est.numel() // (2 * F)
    tok = _pb("decompress_rows", 0.0, 4.0 * (est.numel() + z.numel())) if _PROF is not None else None
    check(_lib.load().seb200_decompress_rows(ptr(est), rows, F, ptr(z), z.shape[-1], stream_ptr()), "seb200_decompress_rows")
    _pe(tok)
    return z


def spec_decompress_rows(spec: torch.Tensor, z: torch.Tensor):
    require_cuda(spec)
    sr = torch.view_as_real(spec.contiguous())
    B, F, T = spec.shape
    tok = _pb("decompress_rows", 0.0, 4.0 * (2 * B * F * T + z.numel())) if _PROF is not None else None
    check(_lib.load().seb200_spec_decompress_rows(ptr(sr), B, F, T, ptr(z), z.shape[-1], stream_ptr()), "seb200_spec_decompress_rows")
    _pe(tok)
    return z


def overlap_add(frames: torch.Tensor, B: int, T: int, inv_env: torch.Tensor, c: Optional[torch.Tensor], out: torch.Tensor):
    _f32c(frames, inv_env, c, out)
    Lout = 100 * (T - 1)
    tok = _pb("overlap_add", 0.0, 4.0 * (frames.numel() + out.numel())) if _PROF is not None else None
    check(_lib.load().seb200_overlap_add(ptr(frames), B, T, frames.shape[-1], ptr(inv_env), ptr(c), ptr(out), Lout,
                                         out.shape[-1], stream_ptr()), "seb200_overlap_add")
    _pe(tok)
    return out


def conv1x1_in3(in3, w, bias, out):
    _f32c(in3, w, bias, out)
    pixels = in3.numel() // 3
    tok = _pb("conv1x1_in3", 2.0 * 3 * 64 * pixels, 4.0 * (in3.numel() + out.numel())) if _PROF is not None else None
    check(_lib.load().seb200_conv1x1_in3(ptr(in3), pixels, ptr(w), ptr(bias), ptr(out), stream_ptr()), "seb200_conv1x1_in3")
    _pe(tok)
    return out


def inorm_workspace(B: int, pix_per_b: int, C_: int, device) -> torch.Tensor:
    nbytes = _lib.load().seb200_inorm_workspace_bytes(B, pix_per_b, C_)
    return torch.empty((nbytes + 7) // 8, device=device, dtype=torch.float64)


def inorm_stats(x, B: int, pix_per_b: int, C_: int, stats, workspace):
    _f32c(x, stats)
    tok = _pb("inorm_stats", 0.0, 4.0 * B * pix_per_b * C_) if _PROF is not None else None
    check(_lib.load().seb200_inorm_stats(ptr(x), B, pix_per_b, C_, ptr(stats), ptr(workspace), workspace.numel() * 8,
                                         stream_ptr()), "seb200_inorm_stats")
    _pe(tok)
    return stats


def inorm_prelu(x, B: int, pix_per_b: int, stats, gamma, beta, slope, y):
    """y fp32 [.., 64] or, when y is bfloat16 [pixels, 2, 64], the pre-split conv-input format (hi | lo per pixel)."""
    _f32c(x, stats, gamma, beta, slope)
    split = y.dtype == torch.bfloat16
    if not split:
        _f32c(y)
    elif not (y.is_contiguous() and y.is_cuda):
        raise RuntimeError("split output must be a contiguous CUDA bfloat16 tensor")
    tok = _pb("inorm_prelu", 0.0, 8.0 * B * pix_per_b * 64) if _PROF is not None else None
    check(_lib.load().seb200_inorm_prelu(ptr(x), B, pix_per_b, 64, ptr(stats), ptr(gamma), ptr(beta), ptr(slope), ptr(y),
                                         int(split), stream_ptr()), "seb200_inorm_prelu")
    _pe(tok)
    return y


def split_planes(x, y):
    """fp32 [pixels, 64] -> bfloat16 [pixels, 2, 64] (hi | lo), hi + lo == x to 2^-17."""
    _f32c(x)
    if y.dtype != torch.bfloat16 or not y.is_contiguous() or not y.is_cuda:
        raise RuntimeError("split output must be a contiguous CUDA bfloat16 tensor")
    tok = _pb("split_planes", 0.0, 8.0 * x.numel()) if _PROF is not None else None
    check(_lib.load().seb200_split_planes(ptr(x), x.numel() // 64, ptr(y), stream_ptr()), "seb200_split_planes")
    _pe(tok)
    return y


def mask_conv(x, rows: int, Fin: int, w, bias: float, out):
    _f32c(x, w, out)
    tok = _pb("mask_conv", 2.0 * 128 * rows * (Fin - 1), 4.0 * (rows * Fin * 64 + rows * (Fin - 1))) if _PROF is not None else None
    check(_lib.load().seb200_mask_conv(ptr(x), rows, Fin, ptr(w), float(bias), ptr(out), stream_ptr()), "seb200_mask_conv")
    _pe(tok)
    return out


def complex_conv(x, B: int, rows_per_b: int, Fin: int, stats, gamma, beta, slope, w, bias, out):
    _f32c(x, stats, gamma, beta, slope, w, bias, out)
    tok = _pb("complex_conv", 2.0 * 256 * B * rows_per_b * (Fin - 1), 4.0 * (B * rows_per_b * Fin * 64 + 2 * B * rows_per_b * (Fin - 1))) if _PROF is not None else None
    check(_lib.load().seb200_complex_conv(ptr(x), B, rows_per_b, Fin, ptr(stats), ptr(gamma), ptr(beta), ptr(slope), ptr(w),
                                          ptr(bias), ptr(out), stream_ptr()), "seb200_complex_conv")
    _pe(tok)
    return out


def mask_recombine(mask_raw, mask_stats, B: int, rows_per_b: int, F: int, scalars, slope_f, in3, cplx, est, mask_out=None):
    _f32c(mask_raw, mask_stats, slope_f, in3, cplx, est, mask_out)
    g, b, s1, wf, bf = (float(v) for v in scalars)
    tok = _pb("mask_recombine", 0.0, 4.0 * 8 * B * rows_per_b * F) if _PROF is not None else None
    check(_lib.load().seb200_mask_recombine(ptr(mask_raw), ptr(mask_stats), B, rows_per_b, F, g, b, s1, wf, bf, ptr(slope_f),
                                            ptr(in3), ptr(cplx), ptr(est), ptr(mask_out), stream_ptr()), "seb200_mask_recombine")
    _pe(tok)
    return est


def split_ri(est, re, im):
    _f32c(est, re, im)
    tok = _pb("split_ri", 0.0, 8.0 * est.numel()) if _PROF is not None else None
    check(_lib.load().seb200_split_ri(ptr(est), est.numel() // 2, ptr(re), ptr(im), stream_ptr()), "seb200_split_ri")
    _pe(tok)


def make_seq(nseq: int, n: int, inner: int, outer_stride: int, pos_stride: int) -> SebSeq:
    s = SebSeq()
    s.nseq, s.n, s.inner, s.outer_stride, s.pos_stride = nseq, n, inner, outer_stride, pos_stride
    return s


_FRAG_ORDER = [0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15]


def pack_rel_pos(rel_pos_emb: torch.Tensor) -> torch.Tensor:
    """rel_pos_emb [1025, 16] -> float16 table in MMA-fragment order (what attention variant 0 reads): the four halfs a
    lane contributes to a B fragment (k = 2t, 2t+1, 2t+8, 2t+9) are adjacent, so it fetches them with one 8-byte load."""
    return rel_pos_emb.to(torch.float16)[:, _FRAG_ORDER].contiguous()


def attention(qkv, rel_pos_emb, seq: SebSeq, out, variant: int = 0, rel_pos_emb_h=None):
    """variant 0 (mma.sync) / 3 (tcgen05): qkv float16 [tokens, 192] with q pre-scaled (EPI_QKV_F16); variant 1: qkv float32,
    unscaled (fp32 SIMT cross-check)."""
    _f32c(rel_pos_emb, out)
    require_cuda(qkv)
    if variant in (0, 3):
        if qkv.dtype != torch.float16 or not qkv.is_contiguous():
            raise RuntimeError("tensor-core attention reads the float16 q|k|v projection")
        if rel_pos_emb_h is None:
            rel_pos_emb_h = pack_rel_pos(rel_pos_emb)
        if rel_pos_emb_h.dtype != torch.float16 or not rel_pos_emb_h.is_contiguous():
            raise RuntimeError("rel_pos_emb_h must be a contiguous float16 copy of the embedding table")
    else:
        _f32c(qkv)
    tok = _pb("attention", 96.0 * 4 * seq.nseq * seq.n * seq.n, (4.0 if variant == 1 else 2.0) * qkv.numel() + 4.0 * out.numel()) if _PROF is not None else None
    check(_lib.load().seb200_attention(ptr(qkv), ptr(rel_pos_emb), ptr(rel_pos_emb_h), C.byref(seq), ptr(out), variant, stream_ptr()), "seb200_attention")
    _pe(tok)
    return out


def dwconv_bn_swish(x, seq: SebSeq, w, bn_scale, bn_shift, y):
    """x, y [tokens, 128]: float32 -> float32, float32 -> float16 (the model's default: half the write) or float16 -> float16; fp32 arithmetic inside"""
    _f32c(w, bn_scale, bn_shift)
    if y.dtype == torch.float16:
        if x.dtype not in (torch.float16, torch.float32) or not (x.is_contiguous() and y.is_contiguous() and x.is_cuda and y.is_cuda):
            raise RuntimeError("dwconv_bn_swish: contiguous CUDA tensors expected (x float32 or float16, y float16)")
        xh = x.dtype == torch.float16
        tok = _pb("dwconv", 62.0 * x.numel(), (4.0 if xh else 6.0) * x.numel()) if _PROF is not None else None
        check(_lib.load().seb200_dwconv_bn_swish_f16(ptr(x), int(xh), C.byref(seq), ptr(w), ptr(bn_scale), ptr(bn_shift), ptr(y), stream_ptr()),
              "seb200_dwconv_bn_swish_f16")
        _pe(tok)
        return y
    _f32c(x, y)
    tok = _pb("dwconv", 62.0 * x.numel(), 8.0 * x.numel()) if _PROF is not None else None
    check(_lib.load().seb200_dwconv_bn_swish(ptr(x), C.byref(seq), ptr(w), ptr(bn_scale), ptr(bn_shift), ptr(y), stream_ptr()),
          "seb200_dwconv_bn_swish")
    _pe(tok)
    return y


def layernorm_residual(x, gamma, beta, resid, out):
    _f32c(x, gamma, beta, resid, out)
    tok = _pb("layernorm_residual", 0.0, 12.0 * x.numel()) if _PROF is not None else None
    check(_lib.load().seb200_layernorm_residual(ptr(x), x.numel() // 64, ptr(gamma), ptr(beta), ptr(resid), ptr(out), stream_ptr()),
          "seb200_layernorm_residual")
    _pe(tok)
    return out
