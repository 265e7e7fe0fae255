"""ctypes binding of libseb200.so (include/seb200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc
(``build.py``); if that fails, or a call returns non-zero, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libseb200.so")
if os.environ.get("SEB200_LIB_SUFFIX"):      # development switch: experiment builds of tools/build_variant_lib.sh
    _LIB_PATH = os.path.join(_HERE, "libseb200_" + os.environ["SEB200_LIB_SUFFIX"] + ".so")

LOAD_ROWS, LOAD_ROWS_LN, LOAD_CONV, LOAD_HANKEL, LOAD_CONV_SPLIT, LOAD_ROWS2, LOAD_CONV_ADJ, LOAD_ROWS_F16 = 0, 1, 2, 3, 4, 5, 6, 7
EPI_BIAS, EPI_SWISH, EPI_GLU, EPI_RESID, EPI_SUBPIXEL, EPI_COMPRESS, EPI_QKV_F16, EPI_GATE, EPI_RESID_SCALE, EPI_GLU_F16 = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
ENGINE_TCGEN05, ENGINE_SIMT, ENGINE_TCGEN05_F32 = 0, 1, 2
ABI_VERSION = 2          # SEB200_ABI_VERSION in include/seb200.h

_fp = C.c_void_p  # raw device pointers travel as void*


class SebGemm(C.Structure):
    _fields_ = [
        ("loader", C.c_int), ("epilogue", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("a", _fp * 4), ("lda", C.c_longlong),
        ("ln_gamma", _fp), ("ln_beta", _fp),
        ("B", C.c_int), ("T", C.c_int), ("Fin", C.c_int), ("Fout", C.c_int),
        ("taps_t", C.c_int), ("dil", C.c_int), ("stride_f", C.c_int), ("nslots", C.c_int),
        ("w_tc", _fp), ("tc_ntile", C.c_int), ("tc_ntiles", C.c_int), ("tc_planes", C.c_int),
        ("w_simt", _fp), ("simt_npad", C.c_int),
        ("bias", _fp),
        ("out", _fp), ("ldo", C.c_longlong),
        ("resid", _fp), ("ldr", C.c_longlong), ("alpha", C.c_float),
    ]


class SebFfn(C.Structure):
    _fields_ = [("x", _fp), ("out", _fp), ("tokens", C.c_longlong), ("ln_gamma", _fp), ("ln_beta", _fp),
                ("w1_tc", _fp), ("b1", _fp), ("w2_tc", _fp), ("b2", _fp), ("alpha", C.c_float),
                ("post_gamma", _fp), ("post_beta", _fp), ("resid2", _fp)]


class SebSeq(C.Structure):
    _fields_ = [("nseq", C.c_int), ("n", C.c_int), ("inner", C.c_int),
                ("outer_stride", C.c_longlong), ("pos_stride", C.c_longlong)]


class SebPackJob(C.Structure):
    """one job of seb200_pack_weights_device_batch (include/seb200.h): device pointers + the index map of seb200_pack_weights_device"""
    _fields_ = [("w", C.c_void_p), ("w_tc", C.c_void_p), ("w_simt", C.c_void_p),
                ("sn", C.c_longlong), ("s0", C.c_longlong), ("s1", C.c_longlong),
                ("N", C.c_int), ("K", C.c_int), ("n1", C.c_int), ("tc_ntile", C.c_int), ("planes", C.c_int), ("reserved", C.c_int)]


_SIGS = {
    "seb200_gemm": [C.POINTER(SebGemm), C.c_int, _fp],
    "seb200_ffn_fused": [C.POINTER(SebFfn), _fp],
    "seb200_rms_pad": [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp],
    "seb200_scale_pad": [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp],
    "seb200_spec_to_in3": [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp],
    "seb200_in3_to_spec": [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp],
    "seb200_decompress_rows": [_fp, C.c_int, C.c_int, _fp, C.c_int, _fp],
    "seb200_spec_decompress_rows": [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp],
    "seb200_overlap_add": [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, _fp],
    "seb200_conv1x1_in3": [_fp, C.c_longlong, _fp, _fp, _fp, _fp],
    "seb200_inorm_stats": [_fp, C.c_int, C.c_longlong, C.c_int, _fp, _fp, C.c_longlong, _fp],
    "seb200_inorm_prelu": [_fp, C.c_int, C.c_longlong, C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_int, _fp],
    "seb200_split_planes": [_fp, C.c_longlong, _fp, _fp],
    "seb200_mask_conv": [_fp, C.c_longlong, C.c_int, _fp, C.c_float, _fp, _fp],
    "seb200_complex_conv": [_fp, C.c_int, C.c_longlong, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp],
    "seb200_mask_recombine": [_fp, _fp, C.c_int, C.c_longlong, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                              C.c_float, _fp, _fp, _fp, _fp, _fp, _fp],
    "seb200_split_ri": [_fp, C.c_longlong, _fp, _fp, _fp],
    "seb200_attention": [_fp, _fp, _fp, C.POINTER(SebSeq), _fp, C.c_int, _fp],
    "seb200_dwconv_bn_swish": [_fp, C.POINTER(SebSeq), _fp, _fp, _fp, _fp, _fp],
    "seb200_dwconv_bn_swish_f16": [_fp, C.c_int, C.POINTER(SebSeq), _fp, _fp, _fp, _fp, _fp],
    "seb200_layernorm_residual": [_fp, C.c_longlong, _fp, _fp, _fp, _fp, _fp],
    "seb200_packed_weight_sizes": [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "seb200_pack_weights": [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp],
    "seb200_compress_backward_rows": [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp],
    "seb200_stft_fold": [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp],
    "seb200_istft_grad_pad": [_fp, _fp, C.c_int, C.c_int, _fp, _fp],
    "seb200_decompress_backward_spec": [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp],
    "seb200_diffusion_update": [_fp, _fp, C.c_longlong, _fp, _fp, C.c_int, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, _fp],
    # ---- training step
    "seb200_pack_weights_device": [_fp, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int, _fp, _fp, _fp],
    "seb200_pack_weights_device_batch": [C.POINTER(SebPackJob), C.c_int, _fp],
    "seb200_wgrad": [C.POINTER(SebGemm), _fp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, _fp, _fp, _fp, C.c_longlong, _fp],
    "seb200_dropout_mask": [_fp, C.c_longlong, C.c_float, C.c_ulonglong, C.c_ulonglong, _fp],
    "seb200_swish_dropout": [_fp, _fp, C.c_float, _fp, C.c_longlong, _fp],
    "seb200_swish_dropout_bwd": [_fp, _fp, C.c_float, _fp, _fp, C.c_longlong, _fp],
    "seb200_dropout_residual": [_fp, _fp, C.c_float, _fp, _fp, C.c_longlong, _fp],
    "seb200_scale_mask": [_fp, _fp, C.c_float, _fp, C.c_longlong, _fp],
    "seb200_glu": [_fp, C.c_longlong, C.c_int, _fp, _fp],
    "seb200_glu_bwd": [_fp, _fp, C.c_longlong, C.c_int, _fp, _fp],
    "seb200_layernorm_bwd": [_fp, _fp, _fp, _fp, _fp, C.c_longlong, _fp, _fp, _fp, _fp],
    "seb200_bn_sums": [_fp, C.c_longlong, _fp, _fp, _fp],
    "seb200_bn_finalize": [_fp, C.c_double, _fp, _fp, _fp, _fp, _fp, C.c_float, C.c_float, _fp, _fp, _fp],
    "seb200_bn_swish": [_fp, C.c_longlong, _fp, _fp, _fp],
    "seb200_bn_swish_bwd_sums": [_fp, _fp, C.c_longlong, _fp, _fp, _fp, _fp, _fp],
    "seb200_bn_swish_bwd_apply": [_fp, _fp, C.c_longlong, _fp, _fp, _fp, _fp, C.c_double, _fp, _fp, _fp, _fp],
    "seb200_dwconv": [_fp, C.POINTER(SebSeq), _fp, _fp, _fp, _fp, _fp],
    "seb200_dwconv_wgrad": [_fp, _fp, C.POINTER(SebSeq), _fp, _fp, _fp, _fp],
    "seb200_inorm_prelu_bwd": [_fp, _fp, C.c_int, C.c_longlong, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_longlong, _fp],
    "seb200_head_conv": [_fp, C.c_longlong, C.c_int, _fp, _fp, C.c_int, _fp, _fp],
    "seb200_head_conv_bwd": [_fp, _fp, C.c_longlong, C.c_int, _fp, C.c_int, _fp, _fp, _fp, _fp, _fp],
    "seb200_mask_recombine_dev": [_fp, _fp, C.c_int, C.c_longlong, C.c_int, C.POINTER(_fp), _fp, _fp, _fp, _fp, _fp],
    "seb200_mask_tail_bwd": [_fp, _fp, C.c_int, C.c_longlong, C.c_int, C.POINTER(_fp), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp],
    "seb200_conv1x1_in3_wgrad": [_fp, _fp, C.c_longlong, _fp, _fp, _fp, _fp],
    "seb200_merge_ri": [_fp, _fp, C.c_longlong, _fp, _fp],
    "seb200_qkv_to_f16": [_fp, C.c_longlong, _fp, _fp],
    "seb200_attention_train_fwd": [_fp, _fp, C.POINTER(SebSeq), _fp, _fp, _fp],
    "seb200_attention_bwd": [_fp, _fp, C.POINTER(SebSeq), C.c_longlong, _fp, _fp, _fp, _fp, _fp, _fp, C.c_longlong, _fp],
    "seb200_diffusion_embed": [_fp, C.c_int, _fp, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp],
}
_LL_FUNCS = {      # size queries returning long long (or int): name -> argtypes
    "seb200_inorm_workspace_bytes": [C.c_int, C.c_longlong, C.c_int],
    "seb200_workspace_bytes": [C.c_int, C.c_int, C.c_int, C.c_int],
    "seb200_train_workspace_floats": [],
    "seb200_wgrad_workspace_floats": [C.c_int, C.c_int, C.c_int],
    "seb200_wgrad_splits": [C.c_int, C.c_int, C.c_int],
    "seb200_inorm_bwd_workspace_doubles": [C.c_int, C.c_longlong, C.c_int],
    "seb200_attention_bwd_workspace_floats": [C.c_longlong],
    "seb200_launch_count": [],
}
EXPORTS = sorted(list(_SIGS) + list(_LL_FUNCS) + ["seb200_version", "seb200_last_error_string"])

_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load(build_if_missing: bool = True):
    """dlopen libseb200.so (building it first if absent) and attach signatures."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if build_if_missing and _build.have_nvcc():
            _build.build()          # no-op when the source digest matches the stamp; rebuilds a stale library after any csrc / header edit
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing and cannot be built here: run `python speech-enhancement_b200/build.py`")
        lib = C.CDLL(_LIB_PATH)
        lib.seb200_version.restype = C.c_int
        if lib.seb200_version() != ABI_VERSION:
            raise RuntimeError(f"{_LIB_PATH} reports ABI version {lib.seb200_version()}, this binding expects {ABI_VERSION}: rebuild "
                               "(`python speech-enhancement_b200/build.py --force`)")
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name, args in _LL_FUNCS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int if name == "seb200_wgrad_splits" else C.c_longlong
        lib.seb200_version.restype = C.c_int
        lib.seb200_last_error_string.restype = C.c_char_p
        _lib = lib
        return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().seb200_last_error_string().decode(errors="replace")
        kind = "argument error" if rc < 0 else "CUDA error"
        raise RuntimeError(f"{what}: {kind} {rc}: {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("se_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device")


def launch_count() -> int:
    return int(load().seb200_launch_count())
