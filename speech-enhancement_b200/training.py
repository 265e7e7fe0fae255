"""Training step of the generator on libseb200 (SURVEY 8f row f1, 8e training; BASELINE configs[4]).

What the reference does (core/function.py:218-277): ``model.train()``; ``est_real, est_imag = model(noisy_spec)`` with dropout
(conformer.py:125,139,141) and BatchNorm1d batch statistics + running-stat update (conformer.py:167; SyncBatchNorm under
main_gan.py:154-155); losses; ``loss.backward()``; DDP's mean all-reduce of the 1,834,833 gradients (main_gan.py:168-171);
``optimizer.step()``.

Here ``TSCNet.forward`` in train mode runs ``GeneratorFunction``: one ``torch.autograd.Function`` around the whole generator.

* forward: the train-mode generator on the CUDA kernels -- GEMM engine in its fp32-grade mode (tcgen05 with three bf16 planes per
  operand, or the fp32 FFMA loop), 3xTF32 attention with the row log-sum-exp kept, explicit dropout masks (Philox, or injected by
  the caller), BatchNorm batch statistics from a two-stage fp64 reduction whose local sums are all-reduced under SyncBatchNorm.
  Every intermediate the backward needs stays in a per-shape buffer set (about 10 GB at 4 x 2 s).
* backward: hand-written kernels for every op (dgrad = the same GEMM engine with transposed weight images / the adjoint-conv loader,
  wgrad = split-M contraction, LayerNorm / BatchNorm / InstanceNorm+PReLU / GLU / Swish / heads / attention with the rel-pos table).
  Parameter gradients are written straight into ONE flat fp32 buffer (parameter order of ``named_parameters()``); the Function returns
  views of it, so ``p.grad`` of all 335 parameters alias one 7.34 MB tensor and the data-parallel exchange is a single NCCL all-reduce
  (``allreduce_gradients``).  torch's DistributedDataParallel also works unchanged (it sees ordinary ``.grad`` tensors).

Why fp32-grade arithmetic: with random (Kaiming) weights this network amplifies a perturbation ~100x on its way to the gradients -- the
reference's own float32 autograd is 2-3e-3 (rel-L2) away from the same graph in float64 (tests/golden/train_b2_L10000.npz) -- so the
two-plane split of the inference path (2^-17 operands) is not enough here.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops, train_ops as T
from ._lib import (EPI_BIAS, EPI_RESID, EPI_SUBPIXEL, LOAD_CONV, LOAD_CONV_ADJ, LOAD_ROWS, LOAD_ROWS_LN)

P_DROP = 0.2            # generator.py:60-65 (attn_dropout = ff_dropout = 0.2)
KEEP_SCALE = 1.0 / (1.0 - P_DROP)
BN_MOMENTUM, BN_EPS = 0.1, 1e-5


def dropout_sites():
    """[(site, channels)] in forward order; the same list synth.dropout_sites() builds for the tests"""
    sites = []
    for i in range(1, 5):
        for ax in ("time", "freq"):
            p = f"TSCB_{i}.{ax}_conformer"
            sites += [(f"{p}.ff1.drop1", 256), (f"{p}.ff1.drop2", 64), (f"{p}.attn.drop", 64), (f"{p}.ff2.drop1", 256), (f"{p}.ff2.drop2", 64)]
    return sites


class TrainState:
    """Per-module training state: flat gradient buffers, device weight images, per-shape activation buffers."""

    def __init__(self, model):
        self.model = model
        self.engine = "tcgen05_f32"          # "tcgen05_f32" (three-plane tcgen05) | "simt" (fp32 FFMA loop)
        self.seed = 0x5EB200
        self.step = 0
        self.injected_masks: Optional[Dict[str, torch.Tensor]] = None      # site -> uint8/bool [B, T, Fh, C] (tests inject the reference's draw)
        self.sync_bn_group = None            # process group for SyncBatchNorm statistics (None: default group when a SyncBatchNorm child exists)
        self._flat = [None, None]
        self._views = [None, None]
        self._names = None
        self._images = None
        self._bufs: Dict[tuple, dict] = {}

    # ---- flat gradient buffers -------------------------------------------------------------------------------------------------------
    def params(self):
        if self._names is None:
            self._names = [n for n, _ in self.model.named_parameters()]
        return self._names, [p for _, p in self.model.named_parameters()]

    def flat(self, which: int, device):
        names, params = self.params()
        if self._flat[which] is None or self._flat[which].device != device:
            total = sum(p.numel() for p in params)
            buf = torch.zeros(total, device=device, dtype=torch.float32)
            views, off = {}, 0
            for n, p in zip(names, params):
                views[n] = buf[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self._flat[which], self._views[which] = buf, views
        return self._flat[which], self._views[which]

    def pick_flat(self, device):
        """the flat buffer this backward writes: never the one the parameters' current .grad tensors alias (autograd then ADDS the
        returned views to them, which is the accumulation semantics of loss.backward())"""
        _, params = self.params()
        alias0 = False
        if self._flat[0] is not None:
            base = self._flat[0].data_ptr()
            end = base + self._flat[0].numel() * 4
            alias0 = any(p.grad is not None and base <= p.grad.data_ptr() < end for p in params)
        return self.flat(1 if alias0 else 0, device)

    def grad_buffer(self) -> Optional[torch.Tensor]:
        """the flat buffer all current .grad tensors alias (None if they do not: e.g. gradients were accumulated elsewhere)"""
        _, params = self.params()
        for which in (0, 1):
            buf = self._flat[which]
            if buf is None:
                continue
            base, off, ok = buf.data_ptr(), 0, True
            for p in params:
                if p.grad is None or p.grad.data_ptr() != base + 4 * off or not p.grad.is_contiguous():
                    ok = False
                    break
                off += p.numel()
            if ok:
                return buf
        return None

    # ---- weight images (rebuilt from the live parameters every forward) -----------------------------------------------------------------
    def images(self, device):
        if self._images is not None and self._images["device"] == device and self._images["engine"] == self.engine:
            return self._images
        tc = self.engine != "simt"
        mk = lambda N, K, nt: T.alloc_packed(N, K, nt, 3, device, tc, not tc)
        W = {"device": device, "engine": self.engine}
        for i in range(1, 5):
            for ax in ("time", "freq"):
                p = f"TSCB_{i}.{ax}_conformer"
                for ff in ("ff1", "ff2"):
                    W[f"{p}.{ff}.w1"], W[f"{p}.{ff}.w1T"] = mk(256, 64, 256), mk(64, 256, 64)
                    W[f"{p}.{ff}.w2"], W[f"{p}.{ff}.w2T"] = mk(64, 256, 64), mk(256, 64, 256)
                W[f"{p}.qkv"], W[f"{p}.qkvT"] = mk(192, 64, 192), mk(64, 192, 64)
                W[f"{p}.out"], W[f"{p}.outT"] = mk(64, 64, 64), mk(64, 64, 64)
                W[f"{p}.pw1"], W[f"{p}.pw1T"] = mk(256, 64, 256), mk(64, 256, 64)
                W[f"{p}.pw2"], W[f"{p}.pw2T"] = mk(64, 128, 64), mk(128, 64, 128)
                W[f"{p}.qkv_cat"] = torch.empty(192, 64, device=device, dtype=torch.float32)
                W[f"{p}.dw"] = torch.empty(31, 128, device=device, dtype=torch.float32)
                W[f"{p}.dw_rev"] = torch.empty(31, 128, device=device, dtype=torch.float32)
        for blk in ("dense_encoder.dilated_dense", "mask_decoder.dense_block", "complex_decoder.dense_block"):
            for i in range(1, 5):
                W[f"{blk}.conv{i}"] = mk(64, 384 * i, 64)
                for j in range(i):
                    W[f"{blk}.conv{i}.adj{j}"] = mk(64, 384, 64)
        W["dense_encoder.conv_2"], W["dense_encoder.conv_2.adj"] = mk(64, 192, 64), mk(64, 192, 64)
        for d in ("mask_decoder", "complex_decoder"):
            W[f"{d}.sub_pixel"], W[f"{d}.sub_pixel.adj"] = mk(128, 192, 128), mk(64, 384, 64)
        W["ones128"] = torch.ones(128, device=device, dtype=torch.float32)
        W["zeros128"] = torch.zeros(128, device=device, dtype=torch.float32)
        self._images = W
        return W

    def pack(self, prm, device):
        """device-side packing of every GEMM weight for this step: forward image + the dgrad image(s)"""
        W = self.images(device)
        jobs = []                                                                            # one batched launch sequence at the end (176 images: 8 launches)
        job = lambda *a, **k: jobs.append(T.pack_job(*a, **k))
        lin = lambda pw, w, N, K: job(pw, w, N, K, K, K, 0, 1)                               # W [N, K]
        linT = lambda pw, w, N, K: job(pw, w, K, N, N, 1, 0, K)                              # W^T [K, N] from W [N, K]
        for i in range(1, 5):
            for ax in ("time", "freq"):
                p = f"TSCB_{i}.{ax}_conformer"
                for ff in ("ff1", "ff2"):
                    w1, w2 = prm[f"{p}.{ff}.fn.fn.net.0.weight"], prm[f"{p}.{ff}.fn.fn.net.3.weight"]
                    lin(W[f"{p}.{ff}.w1"], w1, 256, 64); linT(W[f"{p}.{ff}.w1T"], w1, 256, 64)
                    lin(W[f"{p}.{ff}.w2"], w2, 64, 256); linT(W[f"{p}.{ff}.w2T"], w2, 64, 256)
                cat = W[f"{p}.qkv_cat"]
                cat[:64].copy_(prm[f"{p}.attn.fn.to_q.weight"]); cat[64:].copy_(prm[f"{p}.attn.fn.to_kv.weight"])
                lin(W[f"{p}.qkv"], cat, 192, 64); linT(W[f"{p}.qkvT"], cat, 192, 64)
                wo = prm[f"{p}.attn.fn.to_out.weight"]
                lin(W[f"{p}.out"], wo, 64, 64); linT(W[f"{p}.outT"], wo, 64, 64)
                w1, w2 = prm[f"{p}.conv.net.2.weight"], prm[f"{p}.conv.net.7.weight"]          # (256, 64, 1), (64, 128, 1)
                lin(W[f"{p}.pw1"], w1, 256, 64); linT(W[f"{p}.pw1T"], w1, 256, 64)
                lin(W[f"{p}.pw2"], w2, 64, 128); linT(W[f"{p}.pw2T"], w2, 64, 128)
                dw = prm[f"{p}.conv.net.4.conv.weight"]                                        # (128, 1, 31)
                W[f"{p}.dw"].copy_(dw.squeeze(1).t())
                W[f"{p}.dw_rev"].copy_(dw.squeeze(1).flip(-1).t())
        for blk in ("dense_encoder.dilated_dense", "mask_decoder.dense_block", "complex_decoder.dense_block"):
            for i in range(1, 5):
                w = prm[f"{blk}.conv{i}.weight"]                                               # (64, 64 i, 2, 3)
                cin = 64 * i
                job(W[f"{blk}.conv{i}"], w, 64, 6 * cin, cin, cin * 6, 1, 6)         # K order (tap, cin)
                for j in range(i):                                                             # adjoint of slot j: W'[ci, (tap, co)] = w[co, 64 j + ci, tap]
                    job(W[f"{blk}.conv{i}.adj{j}"], w, 64, 384, 64, 6, 1, cin * 6, w_offset=64 * j * 6)
        w = prm["dense_encoder.conv_2.0.weight"]                                               # (64, 64, 1, 3)
        job(W["dense_encoder.conv_2"], w, 64, 192, 64, 192, 1, 3)
        job(W["dense_encoder.conv_2.adj"], w, 64, 192, 64, 3, 1, 192)
        for d in ("mask_decoder", "complex_decoder"):
            w = prm[f"{d}.sub_pixel.conv.weight"]                                              # (128, 64, 1, 3)
            job(W[f"{d}.sub_pixel"], w, 128, 192, 64, 192, 1, 3)
            job(W[f"{d}.sub_pixel.adj"], w, 64, 384, 128, 3, 1, 192)                  # W'[ci, (kf, co)] = w[co, ci, kf]
        T.pack_device_batch(jobs)
        return W

    # ---- per-shape activation buffers ---------------------------------------------------------------------------------------------------
    def buffers(self, B: int, Tn: int, device) -> dict:
        key = (str(device), B, Tn)
        if key in self._bufs:
            return self._bufs[key]
        F = self.model.num_features
        Fh = (F - 1) // 2 + 1
        P, M = B * Tn * F, B * Tn * Fh
        f32 = dict(device=device, dtype=torch.float32)
        e = lambda *s: torch.empty(*s, **f32)

        def dense(pix):
            return {"raw": [e(pix, 64) for _ in range(4)], "out": [e(pix, 64) for _ in range(4)], "stats": [e(B, 64, 2) for _ in range(4)]}

        S = {"enc": {"raw0": e(P, 64), "stats0": e(B, 64, 2), "a0": e(P, 64), "dense": dense(P), "rawc2": e(M, 64), "statsc2": e(B, 64, 2)},
             "x": [e(M, 64) for _ in range(9)], "conf": {},
             "dec": {d: {"dense": dense(M), "sp": e(B * Tn * 2 * Fh, 64), "stats": e(B, 64, 2)} for d in ("mask_decoder", "complex_decoder")},
             "act_c": e(B * Tn * 2 * Fh, 64), "mask_raw": e(B * Tn, F), "stats1": e(B, 1, 2), "cplx": e(B * Tn, F, 2), "est": e(B * Tn, F, 2),
             "in_ws": ops.inorm_workspace(B, Tn * 2 * Fh, 64, device)}
        for i in range(1, 5):
            for ax in ("time", "freq"):
                S["conf"][f"TSCB_{i}.{ax}_conformer"] = {
                    "a1": e(M, 256), "h1": e(M, 256), "y1": e(M, 64), "qkv": e(M, 192), "o": e(M, 64), "lse": e(M, 4), "y2": e(M, 64),
                    "a3": e(M, 256), "u": e(M, 128), "c": e(M, 128), "v": e(M, 128), "y3": e(M, 64), "a4": e(M, 256), "h4": e(M, 256), "y4": e(M, 64),
                    "bn_ss": e(256), "bn_mr": e(256), "bn_sums": torch.empty(256, device=device, dtype=torch.float64),
                    "masks": {k: torch.empty(M, c, device=device, dtype=torch.uint8) for k, c in (("ff1.drop1", 256), ("ff1.drop2", 64), ("attn.drop", 64),
                                                                                                 ("ff2.drop1", 256), ("ff2.drop2", 64))}}
        # gradient scratch
        S["g"] = {"G": e(M, 64), "H": e(M, 64), "t64": e(M, 64), "t64b": e(M, 64), "g256": e(M, 256), "g192": e(M, 192), "g128a": e(M, 128), "g128b": e(M, 128),
                  "tmp64": e(M, 64), "pix": [e(P, 64) for _ in range(6)], "sp": e(B * Tn * 2 * Fh, 64), "dest": e(B * Tn, F, 2), "dp1": e(B * Tn, F),
                  "dmraw": e(B * Tn, F), "bn_sums": torch.empty(256, device=device, dtype=torch.float64)}
        if len(self._bufs) > 1:
            self._bufs.clear()
        self._bufs[key] = S
        return S


def _state(model) -> TrainState:
    st = model.__dict__.get("_train_state")
    if st is None:
        st = TrainState(model)
        model.__dict__["_train_state"] = st
    return st


# =====================================================================================================================================
# forward / backward of the pieces
# =====================================================================================================================================
class _Ctx:
    """everything one forward / backward pair shares"""

    def __init__(self, st: TrainState, prm: Dict[str, torch.Tensor], W: dict, S: dict, B: int, Tn: int, F: int, device):
        self.st, self.prm, self.W, self.S, self.B, self.T, self.F, self.dev = st, prm, W, S, B, Tn, F, device
        self.Fh = (F - 1) // 2 + 1
        self.M = B * Tn * self.Fh
        self.eng = st.engine
        self.seq_t = ops.make_seq(B * self.Fh, Tn, self.Fh, Tn * self.Fh, self.Fh)
        self.seq_f = ops.make_seq(B * Tn, self.Fh, 1, self.Fh, 1)
        self.gv = None        # name -> gradient view (backward)

    def gemm(self, **kw):
        return ops.gemm(engine=self.eng, **kw)


def _sync_group(model, st: TrainState):
    """process group for BatchNorm statistics: SyncBatchNorm children (main_gan.py:154 converts the eight BatchNorm1d) in an initialised job"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return None, 1
    bn = model.TSCB_1.time_conformer.conv.net[5]
    if isinstance(bn, nn.SyncBatchNorm):
        grp = st.sync_bn_group or bn.process_group
        return (grp if grp is not None else dist.group.WORLD), dist.get_world_size(grp)
    return None, 1


def _dense_fwd(c: _Ctx, prefix: str, D: dict, x0, B, Tn, F):
    """DilatedDenseNet.forward (generator.py:24-32) keeping every raw conv output and statistics"""
    prm, W = c.prm, c.W
    slots = [x0]
    for i in range(1, 5):
        raw, out, stats = D["raw"][i - 1], D["out"][i - 1], D["stats"][i - 1]
        pw = W[f"{prefix}.conv{i}"]
        pw.bias = prm[f"{prefix}.conv{i}.bias"]
        c.gemm(loader=LOAD_CONV, epilogue=EPI_BIAS, M=B * Tn * F, w=pw, a=slots, out=raw, ldo=64, label="t_dconv",
               conv=dict(B=B, T=Tn, Fin=F, Fout=F, taps_t=2, dil=2 ** (i - 1), stride_f=1, nslots=i))
        ops.inorm_stats(raw, B, Tn * F, 64, stats, c.S["in_ws"])
        ops.inorm_prelu(raw, B, Tn * F, stats, prm[f"{prefix}.norm{i}.weight"], prm[f"{prefix}.norm{i}.bias"], prm[f"{prefix}.prelu{i}.weight"], out)
        slots = [out] + slots
    return D["out"][3]


def _dense_bwd(c: _Ctx, prefix: str, D: dict, x0, g_out4, g_x0, accumulate_x0: bool, B, Tn, F, scratch):
    """backward of the dense block: g_out4 = gradient of its output (overwritten), g_x0 receives (or accumulates) the input's gradient.
    scratch: 4 [pix, 64] buffers (gradients of out_1..out_3 and the raw-output gradient)."""
    prm, W, gv = c.prm, c.W, c.gv
    pix = B * Tn * F
    g_o = {4: g_out4, 3: scratch[0], 2: scratch[1], 1: scratch[2]}
    g_raw = scratch[3]
    written = {3: False, 2: False, 1: False, 0: accumulate_x0}
    outs = {0: x0, 1: D["out"][0], 2: D["out"][1], 3: D["out"][2], 4: D["out"][3]}
    targets = {0: g_x0, 1: g_o[1], 2: g_o[2], 3: g_o[3]}
    for i in range(4, 0, -1):
        raw, stats = D["raw"][i - 1], D["stats"][i - 1]
        T.inorm_prelu_bwd(raw, g_o[i], B, Tn * F, 64, stats, prm[f"{prefix}.norm{i}.weight"], prm[f"{prefix}.norm{i}.bias"], prm[f"{prefix}.prelu{i}.weight"],
                          g_raw, gv[f"{prefix}.norm{i}.weight"], gv[f"{prefix}.norm{i}.bias"], gv[f"{prefix}.prelu{i}.weight"])
        slots = [outs[k] for k in range(i - 1, -1, -1)]           # [out_{i-1}, ..., out_1, x0]
        cin = 64 * i
        T.wgrad(loader=LOAD_CONV, M=pix, K=6 * cin, a=slots, g_out=g_raw, ldg=64, N=64, dw=gv[f"{prefix}.conv{i}.weight"], db=gv[f"{prefix}.conv{i}.bias"],
                index_map=(cin, cin * 6, 1, 6), conv=dict(B=B, T=Tn, Fin=F, Fout=F, taps_t=2, dil=2 ** (i - 1), stride_f=1, nslots=i), label="t_dconv_wgrad")
        for j in range(i):                                        # slot j holds out_{i-1-j} (x0 for j = i - 1)
            k = i - 1 - j
            tgt = targets[k]
            acc = written[k]
            c.gemm(loader=LOAD_CONV_ADJ, epilogue=EPI_RESID if acc else EPI_BIAS, M=pix, w=W[f"{prefix}.conv{i}.adj{j}"], a=[g_raw], lda=64, out=tgt, ldo=64,
                   resid=tgt if acc else None, ldr=64 if acc else 0, alpha=1.0, label="t_dconv_dgrad",
                   conv=dict(B=B, T=Tn, Fin=F, Fout=F, taps_t=2, dil=2 ** (i - 1), stride_f=1, nslots=1))
            written[k] = True
    return g_x0


def _masks_for(c: _Ctx, p: str, CS: dict):
    """dropout keep-masks of one conformer: injected by the caller (tests: the reference's draw) or drawn with Philox"""
    st = c.st
    out = {}
    for site in ("ff1.drop1", "ff1.drop2", "attn.drop", "ff2.drop1", "ff2.drop2"):
        m = CS["masks"][site]
        if st.injected_masks is not None:
            src = st.injected_masks[f"{p}.{site}"]
            m.copy_(src.reshape(m.shape).to(device=m.device, dtype=torch.uint8))
        else:
            idx = [s for s, _ in dropout_sites()].index(f"{p}.{site}")
            T.dropout_mask(m, P_DROP, st.seed + 7919 * idx, st.step << 32)
        out[site] = m
    return out


def _ff_fwd(c: _Ctx, p: str, ff: str, x, a_buf, h_buf, y_buf, m1, m2):
    """y = x + 0.5 * drop(W2 drop(swish(W1 LN(x) + b1)) + b2)   (conformer.py:53-71,128-145)"""
    prm, W, M = c.prm, c.W, c.M
    w1, w2 = W[f"{p}.{ff}.w1"], W[f"{p}.{ff}.w2"]
    w1.bias, w2.bias = prm[f"{p}.{ff}.fn.fn.net.0.bias"], prm[f"{p}.{ff}.fn.fn.net.3.bias"]
    c.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_BIAS, M=M, w=w1, a=[x], lda=64, ln=(prm[f"{p}.{ff}.fn.norm.weight"], prm[f"{p}.{ff}.fn.norm.bias"]),
           out=a_buf, ldo=256, label="t_ffn1")
    T.swish_dropout(a_buf, m1, KEEP_SCALE, h_buf)
    t = c.S["g"]["tmp64"]
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=w2, a=[h_buf], lda=256, out=t, ldo=64, label="t_ffn2")
    T.dropout_residual(t, m2, 0.5 * KEEP_SCALE, x, y_buf)
    return y_buf


def _ff_bwd(c: _Ctx, p: str, ff: str, x, a_buf, h_buf, m1, m2, H):
    """H holds the gradient of y on entry and the gradient of x on exit"""
    prm, W, gv, M, g = c.prm, c.W, c.gv, c.M, c.S["g"]
    t64, g256 = g["t64"], g["g256"]
    T.scale_mask(H, m2, 0.5 * KEEP_SCALE, t64)
    T.wgrad(loader=LOAD_ROWS, M=M, K=256, a=[h_buf], lda=256, g_out=t64, ldg=64, N=64, dw=gv[f"{p}.{ff}.fn.fn.net.3.weight"], db=gv[f"{p}.{ff}.fn.fn.net.3.bias"],
            index_map=(256, 256, 0, 1), label="t_ffn_wgrad")
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=W[f"{p}.{ff}.w2T"], a=[t64], lda=64, out=g256, ldo=256, label="t_ffn_dgrad")
    T.swish_dropout_bwd(a_buf, m1, KEEP_SCALE, g256, g256)
    ln = (prm[f"{p}.{ff}.fn.norm.weight"], prm[f"{p}.{ff}.fn.norm.bias"])
    T.wgrad(loader=LOAD_ROWS_LN, M=M, K=64, a=[x], lda=64, ln=ln, g_out=g256, ldg=256, N=256, dw=gv[f"{p}.{ff}.fn.fn.net.0.weight"],
            db=gv[f"{p}.{ff}.fn.fn.net.0.bias"], index_map=(64, 64, 0, 1), label="t_ffn_wgrad")
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=W[f"{p}.{ff}.w1T"], a=[g256], lda=256, out=t64, ldo=64, label="t_ffn_dgrad")
    T.layernorm_bwd(x, ln[0], t64, H, H, gv[f"{p}.{ff}.fn.norm.weight"], gv[f"{p}.{ff}.fn.norm.bias"])


def _conformer_fwd(c: _Ctx, p: str, x_in, x_out, seq, sync):
    """ConformerBlock.forward in train mode + the TSCB outer residual (conformer.py:206-212, generator.py:70,72)"""
    prm, W, M = c.prm, c.W, c.M
    CS = c.S["conf"][p]
    mk = _masks_for(c, p, CS)
    t = c.S["g"]["tmp64"]
    _ff_fwd(c, p, "ff1", x_in, CS["a1"], CS["h1"], CS["y1"], mk["ff1.drop1"], mk["ff1.drop2"])
    # attention
    c.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_BIAS, M=M, w=W[f"{p}.qkv"], a=[CS["y1"]], lda=64, ln=(prm[f"{p}.attn.norm.weight"], prm[f"{p}.attn.norm.bias"]),
           out=CS["qkv"], ldo=192, label="t_qkv")
    T.attention_train_fwd(CS["qkv"], prm[f"{p}.attn.fn.rel_pos_emb.weight"], seq, CS["o"], CS["lse"])
    wo = W[f"{p}.out"]
    wo.bias = prm[f"{p}.attn.fn.to_out.bias"]
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=wo, a=[CS["o"]], lda=64, out=t, ldo=64, label="t_attn_out")
    T.dropout_residual(t, mk["attn.drop"], KEEP_SCALE, CS["y1"], CS["y2"])
    # conv module
    w1 = W[f"{p}.pw1"]
    w1.bias = prm[f"{p}.conv.net.2.bias"]
    c.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_BIAS, M=M, w=w1, a=[CS["y2"]], lda=64, ln=(prm[f"{p}.conv.net.0.weight"], prm[f"{p}.conv.net.0.bias"]),
           out=CS["a3"], ldo=256, label="t_pw1")
    T.glu(CS["a3"], CS["u"])
    T.dwconv(CS["u"], seq, W[f"{p}.dw"], W["ones128"], prm[f"{p}.conv.net.4.conv.bias"], CS["c"])
    T.bn_sums(CS["c"], CS["bn_sums"])
    count = float(M)
    grp, world = sync
    if grp is not None:
        import torch.distributed as dist
        dist.all_reduce(CS["bn_sums"], group=grp)                 # SyncBatchNorm: global sum c, sum c^2 (main_gan.py:154); equal token counts per rank
        count *= world
    bn = f"{p}.conv.net.5"
    bufs = dict(c.st.model.named_buffers())
    T.bn_finalize(CS["bn_sums"], count, prm[f"{bn}.weight"], prm[f"{bn}.bias"], bufs[f"{bn}.running_mean"], bufs[f"{bn}.running_var"],
                  bufs[f"{bn}.num_batches_tracked"], BN_MOMENTUM, BN_EPS, CS["bn_ss"], CS["bn_mr"])
    T.bn_swish(CS["c"], CS["bn_ss"], CS["v"])
    w2 = W[f"{p}.pw2"]
    w2.bias = prm[f"{p}.conv.net.7.bias"]
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=w2, a=[CS["v"]], lda=128, out=CS["y3"], ldo=64, resid=CS["y2"], ldr=64, alpha=1.0, label="t_pw2")
    _ff_fwd(c, p, "ff2", CS["y3"], CS["a4"], CS["h4"], CS["y4"], mk["ff2.drop1"], mk["ff2.drop2"])
    ops.layernorm_residual(CS["y4"], prm[f"{p}.post_norm.weight"], prm[f"{p}.post_norm.bias"], x_in, x_out)
    return x_out


def _conformer_bwd(c: _Ctx, p: str, x_in, seq, sync):
    """g['G'] holds the gradient of the conformer's output on entry; the gradient of its input is left in g['G'] (buffers swapped)"""
    prm, W, gv, M, g = c.prm, c.W, c.gv, c.M, c.S["g"]
    CS = c.S["conf"][p]
    mk = CS["masks"]
    G, H, t64, t64b, g256, g192, g128a, g128b = g["G"], g["H"], g["t64"], g["t64b"], g["g256"], g["g192"], g["g128a"], g["g128b"]
    # post_norm (the outer residual adds G to the input's gradient at the end)
    T.layernorm_bwd(CS["y4"], prm[f"{p}.post_norm.weight"], G, None, H, gv[f"{p}.post_norm.weight"], gv[f"{p}.post_norm.bias"])
    _ff_bwd(c, p, "ff2", CS["y3"], CS["a4"], CS["h4"], mk["ff2.drop1"], mk["ff2.drop2"], H)                      # H = d y3
    # conv module
    T.wgrad(loader=LOAD_ROWS, M=M, K=128, a=[CS["v"]], lda=128, g_out=H, ldg=64, N=64, dw=gv[f"{p}.conv.net.7.weight"], db=gv[f"{p}.conv.net.7.bias"],
            index_map=(128, 128, 0, 1), label="t_pw_wgrad")
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=W[f"{p}.pw2T"], a=[H], lda=64, out=g128a, ldo=128, label="t_pw_dgrad")
    T.bn_swish_bwd_sums(CS["c"], g128a, CS["bn_ss"], CS["bn_mr"], g["bn_sums"])
    count = float(M)
    grp, world = sync
    local = None
    if grp is not None:
        import torch.distributed as dist
        local = g["bn_sums"].clone()                              # SyncBatchNorm returns THIS rank's dgamma / dbeta; dx uses the global sums
        dist.all_reduce(g["bn_sums"], group=grp)
        count *= world
    bn = f"{p}.conv.net.5"
    T.bn_swish_bwd_apply(CS["c"], g128a, CS["bn_ss"], CS["bn_mr"], g["bn_sums"], count, g128b, gv[f"{bn}.weight"], gv[f"{bn}.bias"], local)
    T.dwconv_wgrad(CS["u"], g128b, seq, gv[f"{p}.conv.net.4.conv.weight"], gv[f"{p}.conv.net.4.conv.bias"])
    T.dwconv(g128b, seq, W[f"{p}.dw_rev"], W["ones128"], W["zeros128"], g128a)
    T.glu_bwd(CS["a3"], g128a, g256)
    ln = (prm[f"{p}.conv.net.0.weight"], prm[f"{p}.conv.net.0.bias"])
    T.wgrad(loader=LOAD_ROWS_LN, M=M, K=64, a=[CS["y2"]], lda=64, ln=ln, g_out=g256, ldg=256, N=256, dw=gv[f"{p}.conv.net.2.weight"], db=gv[f"{p}.conv.net.2.bias"],
            index_map=(64, 64, 0, 1), label="t_pw_wgrad")
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=W[f"{p}.pw1T"], a=[g256], lda=256, out=t64, ldo=64, label="t_pw_dgrad")
    T.layernorm_bwd(CS["y2"], ln[0], t64, H, H, gv[f"{p}.conv.net.0.weight"], gv[f"{p}.conv.net.0.bias"])          # H = d y2
    # attention
    T.scale_mask(H, mk["attn.drop"], KEEP_SCALE, t64)
    T.wgrad(loader=LOAD_ROWS, M=M, K=64, a=[CS["o"]], lda=64, g_out=t64, ldg=64, N=64, dw=gv[f"{p}.attn.fn.to_out.weight"], db=gv[f"{p}.attn.fn.to_out.bias"],
            index_map=(64, 64, 0, 1), label="t_attn_wgrad")
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=W[f"{p}.outT"], a=[t64], lda=64, out=t64b, ldo=64, label="t_attn_dgrad")
    T.attention_bwd(CS["qkv"], prm[f"{p}.attn.fn.rel_pos_emb.weight"], seq, CS["o"], CS["lse"], t64b, g192, gv[f"{p}.attn.fn.rel_pos_emb.weight"])
    ln = (prm[f"{p}.attn.norm.weight"], prm[f"{p}.attn.norm.bias"])
    # to_q.weight (64 x 64) and to_kv.weight (128 x 64) are adjacent in the flat buffer: one [192, 64] destination
    dwq = gv[f"{p}.attn.fn.to_q.weight"]
    dqkv_w = torch.as_strided(dwq, (192, 64), (64, 1))
    T.wgrad(loader=LOAD_ROWS_LN, M=M, K=64, a=[CS["y1"]], lda=64, ln=ln, g_out=g192, ldg=192, N=192, dw=dqkv_w, db=None, index_map=(64, 64, 0, 1), label="t_attn_wgrad")
    c.gemm(loader=LOAD_ROWS, epilogue=EPI_BIAS, M=M, w=W[f"{p}.qkvT"], a=[g192], lda=192, out=t64, ldo=64, label="t_attn_dgrad")
    T.layernorm_bwd(CS["y1"], ln[0], t64, H, H, gv[f"{p}.attn.norm.weight"], gv[f"{p}.attn.norm.bias"])            # H = d y1
    _ff_bwd(c, p, "ff1", x_in, CS["a1"], CS["h1"], mk["ff1.drop1"], mk["ff1.drop2"], H)                         # H = d x_in (inner path)
    T.dropout_residual(G, None, 1.0, H, H)                                                                      # + the outer residual
    g["G"], g["H"] = H, G


def generator_forward(model, in3: torch.Tensor):
    """train-mode TSCNet on in3 [B, T, F, 3] -> (est [B*T, F, 2], ctx)"""
    st = _state(model)
    B, Tn, F, _ = in3.shape
    dev = in3.device
    prm = {n: p.detach() for n, p in model.named_parameters()}
    for n, p in prm.items():
        if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
            raise RuntimeError(f"training needs contiguous CUDA float32 parameters ({n})")
    W = st.pack(prm, dev)
    S = st.buffers(B, Tn, dev)
    c = _Ctx(st, prm, W, S, B, Tn, F, dev)
    Fh, M = c.Fh, c.M
    sync = _sync_group(model, st)
    # ---- encoder (generator.py:50-54)
    E = S["enc"]
    e = "dense_encoder"
    ops.conv1x1_in3(in3, prm[f"{e}.conv_1.0.weight"].view(64, 3), prm[f"{e}.conv_1.0.bias"], E["raw0"])
    ops.inorm_stats(E["raw0"], B, Tn * F, 64, E["stats0"], S["in_ws"])
    ops.inorm_prelu(E["raw0"], B, Tn * F, E["stats0"], prm[f"{e}.conv_1.1.weight"], prm[f"{e}.conv_1.1.bias"], prm[f"{e}.conv_1.2.weight"], E["a0"])
    d4 = _dense_fwd(c, f"{e}.dilated_dense", E["dense"], E["a0"], B, Tn, F)
    pw = W[f"{e}.conv_2"]
    pw.bias = prm[f"{e}.conv_2.0.bias"]
    c.gemm(loader=LOAD_CONV, epilogue=EPI_BIAS, M=M, w=pw, a=[d4], out=E["rawc2"], ldo=64, label="t_conv2",
           conv=dict(B=B, T=Tn, Fin=F, Fout=Fh, taps_t=1, dil=1, stride_f=2, nslots=1))
    ops.inorm_stats(E["rawc2"], B, Tn * Fh, 64, E["statsc2"], S["in_ws"])
    ops.inorm_prelu(E["rawc2"], B, Tn * Fh, E["statsc2"], prm[f"{e}.conv_2.1.weight"], prm[f"{e}.conv_2.1.bias"], prm[f"{e}.conv_2.2.weight"], S["x"][0])
    # ---- 4 x TSCB (generator.py:67-74)
    k = 0
    for i in range(1, 5):
        _conformer_fwd(c, f"TSCB_{i}.time_conformer", S["x"][k], S["x"][k + 1], c.seq_t, sync)
        _conformer_fwd(c, f"TSCB_{i}.freq_conformer", S["x"][k + 1], S["x"][k + 2], c.seq_f, sync)
        k += 2
    x = S["x"][8]
    # ---- decoders (generator.py:106-129) + recombination (:158-165)
    for d in ("mask_decoder", "complex_decoder"):
        DD = S["dec"][d]
        d4 = _dense_fwd(c, f"{d}.dense_block", DD["dense"], x, B, Tn, Fh)
        pw = W[f"{d}.sub_pixel"]
        pw.bias = prm[f"{d}.sub_pixel.conv.bias"]
        c.gemm(loader=LOAD_CONV, epilogue=EPI_SUBPIXEL, M=M, w=pw, a=[d4], out=DD["sp"], ldo=64, label="t_subpixel",
               conv=dict(B=B, T=Tn, Fin=Fh, Fout=Fh, taps_t=1, dil=1, stride_f=1, nslots=1))
    m, cd = "mask_decoder", "complex_decoder"
    T.head_conv(S["dec"][m]["sp"], B * Tn, 2 * Fh, prm[f"{m}.conv_1.weight"], prm[f"{m}.conv_1.bias"], 1, S["mask_raw"])
    ops.inorm_stats(S["mask_raw"], B, Tn * F, 1, S["stats1"], S["in_ws"])
    DC = S["dec"][cd]
    ops.inorm_stats(DC["sp"], B, Tn * 2 * Fh, 64, DC["stats"], S["in_ws"])
    ops.inorm_prelu(DC["sp"], B, Tn * 2 * Fh, DC["stats"], prm[f"{cd}.norm.weight"], prm[f"{cd}.norm.bias"], prm[f"{cd}.prelu.weight"], S["act_c"])
    T.head_conv(S["act_c"], B * Tn, 2 * Fh, prm[f"{cd}.conv.weight"], prm[f"{cd}.conv.bias"], 2, S["cplx"])
    scal = [prm[f"{m}.norm.weight"], prm[f"{m}.norm.bias"], prm[f"{m}.prelu.weight"], prm[f"{m}.final_conv.weight"], prm[f"{m}.final_conv.bias"]]
    T.mask_recombine_dev(S["mask_raw"], S["stats1"], B, Tn, F, scal, prm[f"{m}.prelu_out.weight"], in3, S["cplx"], S["est"])
    st.step += 1
    return S["est"], c


def generator_backward(c: _Ctx, in3: torch.Tensor, g_real: torch.Tensor, g_imag: torch.Tensor, flat_views: Dict[str, torch.Tensor]):
    """gradients of every parameter into flat_views from (d final_real, d final_imag), each (B, 1, T, F)"""
    S, prm, W, B, Tn, F, Fh, M = c.S, c.prm, c.W, c.B, c.T, c.F, c.Fh, c.M
    c.gv = gv = flat_views
    g = S["g"]
    sync = _sync_group(c.st.model, c.st)
    dest = T.merge_ri(g_real.contiguous(), g_imag.contiguous(), g["dest"])
    m, cd = "mask_decoder", "complex_decoder"
    gx = g["G"]          # gradient of the TSCB output, then of every conformer input in turn
    pixbuf = g["pix"]    # six [P, 64] scratch tensors; the decoders (F' = 101) use their first halves
    half = lambda t: t.view(-1)[:M * 64].view(M, 64)
    # ---- complex decoder
    DC = S["dec"][cd]
    T.head_conv_bwd(S["act_c"], dest, B * Tn, 2 * Fh, prm[f"{cd}.conv.weight"], 2, g["sp"], gv[f"{cd}.conv.weight"], gv[f"{cd}.conv.bias"])
    T.inorm_prelu_bwd(DC["sp"], g["sp"], B, Tn * 2 * Fh, 64, DC["stats"], prm[f"{cd}.norm.weight"], prm[f"{cd}.norm.bias"], prm[f"{cd}.prelu.weight"],
                      g["sp"], gv[f"{cd}.norm.weight"], gv[f"{cd}.norm.bias"], gv[f"{cd}.prelu.weight"])
    _subpixel_bwd(c, cd, DC, g["sp"], half(pixbuf[4]))
    _dense_bwd(c, f"{cd}.dense_block", DC["dense"], S["x"][8], half(pixbuf[4]), gx, False, B, Tn, Fh, [half(t) for t in pixbuf[:4]])
    # ---- mask decoder
    DM = S["dec"][m]
    scal = [prm[f"{m}.norm.weight"], prm[f"{m}.norm.bias"], prm[f"{m}.prelu.weight"], prm[f"{m}.final_conv.weight"], prm[f"{m}.final_conv.bias"]]
    T.mask_tail_bwd(S["mask_raw"], S["stats1"], B, Tn, F, scal, prm[f"{m}.prelu_out.weight"], in3, dest, g["dp1"], gv[f"{m}.prelu_out.weight"],
                    gv[f"{m}.final_conv.weight"].view(-1), gv[f"{m}.final_conv.bias"])
    T.inorm_prelu_bwd(S["mask_raw"], g["dp1"], B, Tn * F, 1, S["stats1"], prm[f"{m}.norm.weight"], prm[f"{m}.norm.bias"], prm[f"{m}.prelu.weight"],
                      g["dmraw"], gv[f"{m}.norm.weight"], gv[f"{m}.norm.bias"], gv[f"{m}.prelu.weight"])
    T.head_conv_bwd(DM["sp"], g["dmraw"], B * Tn, 2 * Fh, prm[f"{m}.conv_1.weight"], 1, g["sp"], gv[f"{m}.conv_1.weight"], gv[f"{m}.conv_1.bias"])
    _subpixel_bwd(c, m, DM, g["sp"], half(pixbuf[4]))
    _dense_bwd(c, f"{m}.dense_block", DM["dense"], S["x"][8], half(pixbuf[4]), gx, True, B, Tn, Fh, [half(t) for t in pixbuf[:4]])
    # ---- 4 x TSCB
    k = 8
    for i in range(4, 0, -1):
        _conformer_bwd(c, f"TSCB_{i}.freq_conformer", S["x"][k - 1], c.seq_f, sync)
        _conformer_bwd(c, f"TSCB_{i}.time_conformer", S["x"][k - 2], c.seq_t, sync)
        k -= 2
    gx = g["G"]
    # ---- encoder
    E = S["enc"]
    e = "dense_encoder"
    T.inorm_prelu_bwd(E["rawc2"], gx, B, Tn * Fh, 64, E["statsc2"], prm[f"{e}.conv_2.1.weight"], prm[f"{e}.conv_2.1.bias"], prm[f"{e}.conv_2.2.weight"],
                      g["t64"], gv[f"{e}.conv_2.1.weight"], gv[f"{e}.conv_2.1.bias"], gv[f"{e}.conv_2.2.weight"])
    T.wgrad(loader=LOAD_CONV, M=M, K=192, a=[E["dense"]["out"][3]], g_out=g["t64"], ldg=64, N=64, dw=gv[f"{e}.conv_2.0.weight"], db=gv[f"{e}.conv_2.0.bias"],
            index_map=(64, 192, 1, 3), conv=dict(B=B, T=Tn, Fin=F, Fout=Fh, taps_t=1, dil=1, stride_f=2, nslots=1), label="t_conv2_wgrad")
    c.gemm(loader=LOAD_CONV_ADJ, epilogue=EPI_BIAS, M=B * Tn * F, w=W[f"{e}.conv_2.adj"], a=[g["t64"]], lda=64, out=pixbuf[4], ldo=64, label="t_conv2_dgrad",
           conv=dict(B=B, T=Tn, Fin=Fh, Fout=F, taps_t=1, dil=1, stride_f=2, nslots=1))
    _dense_bwd(c, f"{e}.dilated_dense", E["dense"], E["a0"], pixbuf[4], pixbuf[5], False, B, Tn, F, pixbuf[:4])
    T.inorm_prelu_bwd(E["raw0"], pixbuf[5], B, Tn * F, 64, E["stats0"], prm[f"{e}.conv_1.1.weight"], prm[f"{e}.conv_1.1.bias"], prm[f"{e}.conv_1.2.weight"],
                      pixbuf[4], gv[f"{e}.conv_1.1.weight"], gv[f"{e}.conv_1.1.bias"], gv[f"{e}.conv_1.2.weight"])
    T.conv1x1_in3_wgrad(in3, pixbuf[4], gv[f"{e}.conv_1.0.weight"].view(64, 3), gv[f"{e}.conv_1.0.bias"])


def _subpixel_bwd(c: _Ctx, d: str, DD: dict, g_sp, g_d4):
    """SPConvTranspose2d backward (generator.py:85-92): the interleaved output [B*T*2F', 64] is the [B*T*F', 128] matrix of the conv's
    128 output channels, so its gradient is read in place"""
    B, Tn, Fh, M, gv = c.B, c.T, c.Fh, c.M, c.gv
    g128 = g_sp.view(M, 128)
    T.wgrad(loader=LOAD_CONV, M=M, K=192, a=[DD["dense"]["out"][3]], g_out=g128, ldg=128, N=128, dw=gv[f"{d}.sub_pixel.conv.weight"], db=gv[f"{d}.sub_pixel.conv.bias"],
            index_map=(64, 192, 1, 3), conv=dict(B=B, T=Tn, Fin=Fh, Fout=Fh, taps_t=1, dil=1, stride_f=1, nslots=1), label="t_subpixel_wgrad")
    c.gemm(loader=LOAD_CONV_ADJ, epilogue=EPI_BIAS, M=M, w=c.W[f"{d}.sub_pixel.adj"], a=[g128], lda=128, out=g_d4, ldo=64, label="t_subpixel_dgrad",
           conv=dict(B=B, T=Tn, Fin=Fh, Fout=Fh, taps_t=1, dil=1, stride_f=1, nslots=2))


# =====================================================================================================================================
# autograd boundary
# =====================================================================================================================================
class GeneratorFunction(torch.autograd.Function):
    """(in3, *parameters) -> (final_real, final_imag); backward returns views of the flat gradient buffer for the parameters"""

    @staticmethod
    def forward(ctx, model, in3, *params):
        with torch.cuda.device(in3.device):
            est, c = generator_forward(model, in3)
            B, Tn, F, _ = in3.shape
            fr = torch.empty(B, 1, Tn, F, device=in3.device, dtype=torch.float32)
            fi = torch.empty(B, 1, Tn, F, device=in3.device, dtype=torch.float32)
            ops.split_ri(est, fr, fi)
        ctx.c, ctx.in3, ctx.model = c, in3, model
        ctx.set_materialize_grads(True)
        return fr, fi

    @staticmethod
    def backward(ctx, g_real, g_imag):
        c, in3, model = ctx.c, ctx.in3, ctx.model
        st = _state(model)
        with torch.cuda.device(in3.device):
            flat, views = st.pick_flat(in3.device)
            generator_backward(c, in3, g_real.to(torch.float32), g_imag.to(torch.float32), views)
        names, _ = st.params()
        # fresh view objects: autograd adopts an incoming gradient as p.grad (no copy) only when nothing else references the tensor object
        return (None, None) + tuple(views[n].view(views[n].shape) for n in names)


def forward_train(model, x: torch.Tensor):
    """TSCNet.forward in train mode: complex64 (B, F, T) -> (final_real, final_imag) attached to the autograd graph of the parameters"""
    with torch.cuda.device(x.device):
        in3 = ops.spec_to_in3(x.detach().to(torch.complex64))
    params = [p for _, p in model.named_parameters()]
    return GeneratorFunction.apply(model, in3, *params)


def allreduce_gradients(model, group=None, average: bool = True):
    """The data-parallel exchange of the generator step (main_gan.py:168-171 does it with DistributedDataParallel buckets): ONE all-reduce of
    the flat 7.34 MB gradient buffer that every .grad aliases; falls back to one coalesced call over the .grad tensors otherwise."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    world = dist.get_world_size(group)
    buf = _state(model).grad_buffer()
    if buf is not None:
        work = dist.all_reduce(buf, group=group, async_op=True)
        work.wait()
        if average:
            buf.mul_(1.0 / world)
        return buf
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat.mul_(1.0 / world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat
