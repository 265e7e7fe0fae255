"""Tensor-level wrappers over the training entry points of the C ABI (include/seb200.h, "training step"): one function per
exported kernel, same conventions as ``ops.py`` (CUDA fp32 contiguous tensors, raw pointers + the current stream, non-zero
return code -> RuntimeError).  SURVEY 8f row f1: train-mode forward pieces and the backward of every op of the generator."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib, ops
from ._lib import SebGemm, SebSeq, check, ptr, stream_ptr
from .ops import _f32c, _pb, _pe
from .packing import PackedWeight

_ws_cache = {}


def workspace(device, floats: Optional[int] = None) -> torch.Tensor:
    """scratch for the two-stage reductions (per device; sized once, grows on demand)"""
    need = int(floats or _lib.load().seb200_train_workspace_floats())
    key = str(device)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = _ws_cache[key] = torch.empty(max(need, int(_lib.load().seb200_train_workspace_floats())), device=device, dtype=torch.float32)
    return ws


def workspace64(device, doubles: int) -> torch.Tensor:
    key = ("f64", str(device))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < doubles:
        ws = _ws_cache[key] = torch.empty(max(int(doubles), 1 << 20), device=device, dtype=torch.float64)
    return ws


def _mask_ptr(mask):
    if mask is None:
        return 0
    if mask.dtype not in (torch.uint8, torch.bool) or not mask.is_contiguous() or not mask.is_cuda:
        raise RuntimeError("dropout masks are contiguous CUDA uint8 / bool tensors (1 = keep)")
    return mask.data_ptr()


# ---- weight images on the device ---------------------------------------------------------------------------------------------------
def packed_sizes(N: int, K: int, tc_ntile: int, planes: int):
    tcb, sf, kp, ntiles, npad = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_int(), C.c_int()
    check(_lib.load().seb200_packed_weight_sizes(N, K, tc_ntile, planes, C.byref(tcb), C.byref(sf), C.byref(kp), C.byref(ntiles), C.byref(npad)),
          "seb200_packed_weight_sizes")
    return tcb.value, sf.value, kp.value, ntiles.value, npad.value


def alloc_packed(N: int, K: int, tc_ntile: int, planes: int, device, want_tc: bool, want_simt: bool) -> PackedWeight:
    """empty device images for one logical W [N, K] (filled by pack_device every step)"""
    tcb, sf, kp, ntiles, npad = packed_sizes(N, K, tc_ntile, planes)
    w_tc = torch.empty(tcb if want_tc else 0, device=device, dtype=torch.uint8)
    w_simt = torch.empty((kp, npad) if want_simt else (0, npad), device=device, dtype=torch.float32)
    return PackedWeight(N, kp, tc_ntile, ntiles, npad, w_tc, w_simt, None, planes)


def pack_device(pw: PackedWeight, w: torch.Tensor, N: int, K: int, n1: int, sn: int, s0: int, s1: int, w_offset: int = 0):
    """fill pw's images from the live fp32 parameter storage: W[n, k] = w.flatten()[w_offset + n * sn + (k / n1) * s0 + (k % n1) * s1]"""
    if w.dtype != torch.float32 or not w.is_cuda or not w.is_contiguous():
        raise RuntimeError("pack_device reads contiguous CUDA float32 parameters")
    check(_lib.load().seb200_pack_weights_device(w.data_ptr() + 4 * w_offset, N, K, n1, sn, s0, s1, pw.tc_ntile, pw.planes,
                                                 ptr(pw.w_tc) if pw.w_tc.numel() else 0, ptr(pw.w_simt) if pw.w_simt.numel() else 0, stream_ptr()),
          "seb200_pack_weights_device")


def pack_job(pw: PackedWeight, w: torch.Tensor, N: int, K: int, n1: int, sn: int, s0: int, s1: int, w_offset: int = 0):
    """the same request as pack_device as a job descriptor for pack_device_batch"""
    if w.dtype != torch.float32 or not w.is_cuda or not w.is_contiguous():
        raise RuntimeError("pack_device reads contiguous CUDA float32 parameters")
    return _lib.SebPackJob(w.data_ptr() + 4 * w_offset, ptr(pw.w_tc) if pw.w_tc.numel() else None, ptr(pw.w_simt) if pw.w_simt.numel() else None,
                           sn, s0, s1, N, K, n1, pw.tc_ntile, pw.planes, 0)


def pack_device_batch(jobs):
    """every job of the list in ceil(len / 24) launches (a training step re-packs 176 images)"""
    if not jobs:
        return
    arr = (_lib.SebPackJob * len(jobs))(*jobs)
    check(_lib.load().seb200_pack_weights_device_batch(arr, len(jobs), stream_ptr()), "seb200_pack_weights_device_batch")


# ---- weight gradients ---------------------------------------------------------------------------------------------------------------
def wgrad(*, loader: int, M: int, K: int, a: Sequence[torch.Tensor], g_out: torch.Tensor, ldg: int, N: int, dw: torch.Tensor, db: Optional[torch.Tensor],
          index_map, lda: int = 0, ln=None, conv: Optional[dict] = None, k_logical: Optional[int] = None, label: str = "wgrad"):
    """dW[n, k] = sum_m g_out[m, n] * A[m, k] into dw (a view of the flat gradient buffer) through index_map = (n1, sn, s0, s1)"""
    lib = _lib.load()
    _f32c(g_out, dw, db, *a)
    g = SebGemm()
    g.loader, g.M, g.N, g.K = loader, M, N, K
    for i in range(4):
        g.a[i] = ptr(a[i]) if i < len(a) else 0
    g.lda = lda
    if ln is not None:
        g.ln_gamma, g.ln_beta = ptr(ln[0]), ptr(ln[1])
    if conv is not None:
        g.B, g.T, g.Fin, g.Fout = conv["B"], conv["T"], conv["Fin"], conv["Fout"]
        g.taps_t, g.dil, g.stride_f, g.nslots = conv.get("taps_t", 1), conv.get("dil", 1), conv.get("stride_f", 1), conv.get("nslots", 1)
    need = lib.seb200_wgrad_workspace_floats(M, N, K)
    ws = workspace(g_out.device, need)
    n1, sn, s0, s1 = index_map
    tok = _pb(label, 2.0 * M * N * (k_logical or K), 4.0 * M * (N + (k_logical or K))) if ops._PROF is not None else None
    check(lib.seb200_wgrad(C.byref(g), ptr(g_out), ldg, N, k_logical or K, n1, sn, s0, s1, ptr(dw), ptr(db), ptr(ws), ws.numel(), stream_ptr()), "seb200_wgrad")
    _pe(tok)


# ---- elementwise ----------------------------------------------------------------------------------------------------------------------
def dropout_mask(mask: torch.Tensor, p: float, seed: int, offset: int):
    check(_lib.load().seb200_dropout_mask(_mask_ptr(mask), mask.numel(), p, seed, offset, stream_ptr()), "seb200_dropout_mask")
    return mask


def _elem(fn_name, label, nbytes_per, *args):
    tok = _pb(label, 0.0, nbytes_per) if ops._PROF is not None else None
    check(getattr(_lib.load(), fn_name)(*args, stream_ptr()), fn_name)
    _pe(tok)


def swish_dropout(a, mask, scale: float, h):
    _f32c(a, h)
    _elem("seb200_swish_dropout", "swish_dropout", 9.0 * a.numel(), ptr(a), _mask_ptr(mask), scale, ptr(h), a.numel())
    return h


def swish_dropout_bwd(a, mask, scale: float, dh, da):
    _f32c(a, dh, da)
    _elem("seb200_swish_dropout_bwd", "swish_dropout_bwd", 13.0 * a.numel(), ptr(a), _mask_ptr(mask), scale, ptr(dh), ptr(da), a.numel())
    return da


def dropout_residual(t, mask, scale: float, resid, y):
    _f32c(t, resid, y)
    _elem("seb200_dropout_residual", "dropout_residual", 13.0 * t.numel(), ptr(t), _mask_ptr(mask), scale, ptr(resid), ptr(y), t.numel())
    return y


def scale_mask(dy, mask, scale: float, dt):
    _f32c(dy, dt)
    _elem("seb200_scale_mask", "scale_mask", 9.0 * dy.numel(), ptr(dy), _mask_ptr(mask), scale, ptr(dt), dy.numel())
    return dt


def glu(a, u):
    _f32c(a, u)
    M, C2 = a.shape
    _elem("seb200_glu", "glu", 6.0 * a.numel(), ptr(a), M, C2 // 2, ptr(u))
    return u


def glu_bwd(a, du, da):
    _f32c(a, du, da)
    M, C2 = a.shape
    _elem("seb200_glu_bwd", "glu_bwd", 10.0 * a.numel(), ptr(a), ptr(du), M, C2 // 2, ptr(da))
    return da


def layernorm_bwd(x, gamma, dy, add, dx, dgamma, dbeta):
    _f32c(x, gamma, dy, add, dx, dgamma, dbeta)
    ws = workspace(x.device)
    _elem("seb200_layernorm_bwd", "layernorm_bwd", 4.0 * x.numel() * (3 + (add is not None)), ptr(x), ptr(gamma), ptr(dy), ptr(add), ptr(dx), x.numel() // 64,
          ptr(dgamma), ptr(dbeta), ptr(ws))
    return dx


# ---- BatchNorm1d(128), train mode ---------------------------------------------------------------------------------------------------
def bn_sums(c, sums):
    _f32c(c)
    ws = workspace64(c.device, 148 * 4 * 256)
    _elem("seb200_bn_sums", "bn_sums", 4.0 * c.numel(), ptr(c), c.shape[0], ptr(sums), ptr(ws))
    return sums


def bn_finalize(sums, count: float, gamma, beta, running_mean, running_var, nbt, momentum: float, eps: float, scale_shift, mean_rstd):
    check(_lib.load().seb200_bn_finalize(ptr(sums), float(count), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(nbt), momentum, eps,
                                         ptr(scale_shift), ptr(mean_rstd), stream_ptr()), "seb200_bn_finalize")


def bn_swish(c, scale_shift, v):
    _f32c(c, scale_shift, v)
    _elem("seb200_bn_swish", "bn_swish", 8.0 * c.numel(), ptr(c), c.shape[0], ptr(scale_shift), ptr(v))
    return v


def bn_swish_bwd_sums(c, dv, scale_shift, mean_rstd, sums):
    _f32c(c, dv)
    ws = workspace64(c.device, 148 * 4 * 256)
    _elem("seb200_bn_swish_bwd_sums", "bn_bwd_sums", 8.0 * c.numel(), ptr(c), ptr(dv), c.shape[0], ptr(scale_shift), ptr(mean_rstd), ptr(sums), ptr(ws))
    return sums


def bn_swish_bwd_apply(c, dv, scale_shift, mean_rstd, sums, count: float, dc, dgamma, dbeta, sums_local=None):
    _f32c(c, dv, dc)
    _elem("seb200_bn_swish_bwd_apply", "bn_bwd_apply", 12.0 * c.numel(), ptr(c), ptr(dv), c.shape[0], ptr(scale_shift), ptr(mean_rstd), ptr(sums), ptr(sums_local),
          float(count), ptr(dc), ptr(dgamma), ptr(dbeta))
    return dc


def dwconv(x, seq: SebSeq, w, scale, shift, y):
    _f32c(x, w, scale, shift, y)
    tok = _pb("dwconv_raw", 62.0 * x.numel(), 8.0 * x.numel()) if ops._PROF is not None else None
    check(_lib.load().seb200_dwconv(ptr(x), C.byref(seq), ptr(w), ptr(scale), ptr(shift), ptr(y), stream_ptr()), "seb200_dwconv")
    _pe(tok)
    return y


def dwconv_wgrad(u, dc, seq: SebSeq, dw, db):
    _f32c(u, dc, dw, db)
    ws = workspace(u.device)
    tok = _pb("dwconv_wgrad", 62.0 * u.numel(), 8.0 * u.numel()) if ops._PROF is not None else None
    check(_lib.load().seb200_dwconv_wgrad(ptr(u), ptr(dc), C.byref(seq), ptr(dw), ptr(db), ptr(ws), stream_ptr()), "seb200_dwconv_wgrad")
    _pe(tok)


# ---- InstanceNorm2d + PReLU backward ------------------------------------------------------------------------------------------------------
def inorm_prelu_bwd(x, dy, B: int, pix_per_b: int, C_: int, stats, gamma, beta, slope, dx, dgamma, dbeta, dslope):
    _f32c(x, dy, stats, gamma, beta, slope, dx, dgamma, dbeta, dslope)
    lib = _lib.load()
    ws = workspace64(x.device, lib.seb200_inorm_bwd_workspace_doubles(B, pix_per_b, C_))
    tok = _pb("inorm_prelu_bwd", 0.0, 4.0 * 5 * B * pix_per_b * C_) if ops._PROF is not None else None
    check(lib.seb200_inorm_prelu_bwd(ptr(x), ptr(dy), B, pix_per_b, C_, ptr(stats), ptr(gamma), ptr(beta), ptr(slope), ptr(dx), ptr(dgamma), ptr(dbeta),
                                     ptr(dslope), ptr(ws), ws.numel(), stream_ptr()), "seb200_inorm_prelu_bwd")
    _pe(tok)
    return dx


# ---- decoder heads --------------------------------------------------------------------------------------------------------------------
def head_conv(x, rows: int, Fin: int, w, bias, NO: int, out):
    _f32c(x, w, bias, out)
    _elem("seb200_head_conv", "head_conv", 4.0 * rows * Fin * 64, ptr(x), rows, Fin, ptr(w), ptr(bias), NO, ptr(out))
    return out


def head_conv_bwd(x, dout, rows: int, Fin: int, w, NO: int, dx, dw, db):
    _f32c(x, dout, w, dx, dw, db)
    ws = workspace(x.device)
    _elem("seb200_head_conv_bwd", "head_conv_bwd", 8.0 * rows * Fin * 64, ptr(x), ptr(dout), rows, Fin, ptr(w), NO, ptr(dx), ptr(dw), ptr(db), ptr(ws))
    return dx


def _scalar_ptrs(scalars5):
    arr = (C.c_void_p * 5)()
    for i, t in enumerate(scalars5):
        _f32c(t)
        arr[i] = t.data_ptr()
    return arr


def mask_recombine_dev(mask_raw, stats1, B: int, rows_per_b: int, F: int, scalars5, slope_f, in3, cplx, est):
    _f32c(mask_raw, stats1, slope_f, in3, cplx, est)
    arr = _scalar_ptrs(scalars5)
    _elem("seb200_mask_recombine_dev", "mask_recombine", 32.0 * B * rows_per_b * F, ptr(mask_raw), ptr(stats1), B, rows_per_b, F, arr, ptr(slope_f), ptr(in3),
          ptr(cplx), ptr(est))
    return est


def mask_tail_bwd(mask_raw, stats1, B: int, rows_per_b: int, F: int, scalars5, slope_f, in3, dest, dp1, dslope_f, dwf, dbf):
    _f32c(mask_raw, stats1, slope_f, in3, dest, dp1, dslope_f, dwf, dbf)
    arr = _scalar_ptrs(scalars5)
    ws = workspace(mask_raw.device)
    _elem("seb200_mask_tail_bwd", "mask_tail_bwd", 32.0 * B * rows_per_b * F, ptr(mask_raw), ptr(stats1), B, rows_per_b, F, arr, ptr(slope_f), ptr(in3), ptr(dest),
          ptr(dp1), ptr(dslope_f), ptr(dwf), ptr(dbf), ptr(ws))
    return dp1


def conv1x1_in3_wgrad(in3, g, dw, db):
    _f32c(in3, g, dw, db)
    ws = workspace(g.device)
    _elem("seb200_conv1x1_in3_wgrad", "conv1x1_in3_wgrad", 4.0 * g.numel(), ptr(in3), ptr(g), g.numel() // 64, ptr(dw), ptr(db), ptr(ws))


def merge_ri(re, im, est):
    _f32c(re, im, est)
    _elem("seb200_merge_ri", "merge_ri", 16.0 * re.numel(), ptr(re), ptr(im), re.numel(), ptr(est))
    return est


def qkv_to_f16(qkv, out):
    _f32c(qkv)
    _elem("seb200_qkv_to_f16", "qkv_to_f16", 6.0 * qkv.numel(), ptr(qkv), qkv.shape[0], ptr(out))
    return out


# ---- attention ------------------------------------------------------------------------------------------------------------------------
def attention_train_fwd(qkv, rel_pos_emb, seq: SebSeq, out, lse):
    _f32c(qkv, rel_pos_emb, out, lse)
    tok = _pb("attention_train_fwd", 96.0 * 4 * seq.nseq * seq.n * seq.n, 4.0 * (qkv.numel() + out.numel())) if ops._PROF is not None else None
    check(_lib.load().seb200_attention_train_fwd(ptr(qkv), ptr(rel_pos_emb), C.byref(seq), ptr(out), ptr(lse), stream_ptr()), "seb200_attention_train_fwd")
    _pe(tok)
    return out


def attention_bwd(qkv, rel_pos_emb, seq: SebSeq, out, lse, dout, dqkv, drel):
    _f32c(qkv, rel_pos_emb, out, lse, dout, dqkv, drel)
    lib = _lib.load()
    tokens = out.shape[0]
    ws = workspace(qkv.device, lib.seb200_attention_bwd_workspace_floats(tokens))
    tok = _pb("attention_bwd", 2.5 * 96.0 * 4 * seq.nseq * seq.n * seq.n, 4.0 * (2 * qkv.numel() + 2 * out.numel())) if ops._PROF is not None else None
    check(lib.seb200_attention_bwd(ptr(qkv), ptr(rel_pos_emb), C.byref(seq), tokens, ptr(out), ptr(lse), ptr(dout), ptr(dqkv), ptr(drel), ptr(ws), ws.numel(),
                                   stream_ptr()), "seb200_attention_bwd")
    _pe(tok)
    return dqkv
