"""Drop-in ``TSCNet`` for the SCP-GAN / CMGAN generator, running on libseb200 (sm_100a).

Mirrors the reference's module interface for this path
(/root/reference/models/generator.py:132-167):

    TSCNet(num_channel=64, num_features=201).forward(x, diffusion_step=None) -> (final_real, final_imag)

* same constructor arguments, same sub-module tree and therefore the same 359
  ``state_dict`` keys / shapes / dtypes, so ``load_state_dict`` of a reference
  checkpoint (after inference_gan.py:66-68 strips ``module.``) works unchanged;
  ``conv.net.5`` is a real ``nn.BatchNorm1d`` so ``SyncBatchNorm.convert_sync_batchnorm``
  (main_gan.py:154) still finds it.
* ``forward`` takes the complex64 ``(B, 201, T)`` compressed spectrogram and returns two
  fp32 ``(B, 1, T, 201)`` tensors, like the reference.

The parameter-holding sub-modules below are containers only: all arithmetic is done by
the CUDA kernels through ``ops`` in channels-last ``[B, T, F, C]`` fp32.  There is no
CPU path and no PyTorch-eager fallback; non-CUDA input raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import (EPI_BIAS, EPI_GLU, EPI_GLU_F16, EPI_QKV_F16, EPI_RESID, EPI_SUBPIXEL, EPI_SWISH, LOAD_CONV, LOAD_CONV_SPLIT, LOAD_ROWS,
                   LOAD_ROWS_F16, LOAD_ROWS_LN)
from .packing import PackedWeight, conv_weight_matrix, glu_interleave, pack_weight


# ---------------------------------------------------------------------------------------------
# parameter containers (names follow the reference so the state_dict is identical)
# ---------------------------------------------------------------------------------------------
class _Bag(nn.Module):
    """A nameable container; the kernels, not these modules, do the computing."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-modules of the B200 TSCNet are parameter containers; call TSCNet.forward")


def _dense_block(ch: int, depth: int = 4) -> _Bag:
    """DilatedDenseNet parameters (generator.py:6-22)."""
    blk = _Bag()
    for i in range(1, depth + 1):
        blk.add_module(f"conv{i}", nn.Conv2d(ch * i, ch, kernel_size=(2, 3), dilation=(2 ** (i - 1), 1)))
        blk.add_module(f"norm{i}", nn.InstanceNorm2d(ch, affine=True))
        blk.add_module(f"prelu{i}", nn.PReLU(ch))
    return blk


def _feed_forward(dim: int, mult: int, p_drop: float) -> _Bag:
    """Scale(0.5, PreNorm(dim, FeedForward)) parameters (conformer.py:53-71,128-145) -> keys ``fn.fn.net.{0,3}``, ``fn.norm``."""
    ff = _Bag()
    ff.net = nn.Sequential(nn.Linear(dim, dim * mult), nn.Identity(), nn.Dropout(p_drop),
                           nn.Linear(dim * mult, dim), nn.Dropout(p_drop))
    pre = _Bag()
    pre.fn = ff
    pre.norm = nn.LayerNorm(dim)
    scale = _Bag()
    scale.fn = pre
    return scale


def _attention(dim: int, heads: int, dim_head: int, p_drop: float, max_pos: int = 512) -> _Bag:
    """PreNorm(dim, Attention) parameters (conformer.py:74-94)."""
    att = _Bag()
    inner = heads * dim_head
    att.to_q = nn.Linear(dim, inner, bias=False)
    att.to_kv = nn.Linear(dim, inner * 2, bias=False)
    att.to_out = nn.Linear(inner, dim)
    att.rel_pos_emb = nn.Embedding(2 * max_pos + 1, dim_head)
    att.dropout = nn.Dropout(p_drop)
    pre = _Bag()
    pre.fn = att
    pre.norm = nn.LayerNorm(dim)
    return pre


def _conv_module(dim: int, expansion: int, ksize: int) -> _Bag:
    """ConformerConvModule parameters (conformer.py:158-172) -> keys ``net.{0,2,4.conv,5,7}``."""
    inner = dim * expansion
    dw = _Bag()
    dw.conv = nn.Conv1d(inner, inner, ksize, groups=inner)
    mod = _Bag()
    mod.net = nn.Sequential(nn.LayerNorm(dim), nn.Identity(), nn.Conv1d(dim, inner * 2, 1), nn.Identity(), dw,
                            nn.BatchNorm1d(inner), nn.Identity(), nn.Conv1d(inner, dim, 1), nn.Identity(), nn.Dropout(0.0))
    return mod


def _conformer(dim: int) -> _Bag:
    """ConformerBlock(dim, dim_head=dim//4, heads=4, conv_kernel_size=31, dropout 0.2) (generator.py:60-65)."""
    blk = _Bag()
    blk.ff1 = _feed_forward(dim, 4, 0.2)
    blk.attn = _attention(dim, 4, dim // 4, 0.2)
    blk.conv = _conv_module(dim, 2, 31)
    blk.ff2 = _feed_forward(dim, 4, 0.2)
    blk.post_norm = nn.LayerNorm(dim)
    return blk


def _sub_pixel(ch: int) -> _Bag:
    sp = _Bag()
    sp.conv = nn.Conv2d(ch, ch * 2, kernel_size=(1, 3), stride=(1, 1))
    return sp


def _dense_encoder(ch: int) -> _Bag:
    """DenseEncoder(in_channel=3, channels=ch) parameters (generator.py:35-48)."""
    enc = _Bag()
    enc.conv_1 = nn.Sequential(nn.Conv2d(3, ch, (1, 1), (1, 1)), nn.InstanceNorm2d(ch, affine=True), nn.PReLU(ch))
    enc.dilated_dense = _dense_block(ch)
    enc.conv_2 = nn.Sequential(nn.Conv2d(ch, ch, (1, 3), (1, 2), padding=(0, 1)), nn.InstanceNorm2d(ch, affine=True), nn.PReLU(ch))
    return enc


class TSCNet(nn.Module):
    def __init__(self, num_channel: int = 64, num_features: int = 201):
        super().__init__()
        if num_channel != 64:
            raise ValueError("the sm_100a kernels are specialised for num_channel=64 (main_gan.py:145, inference_gan.py:61)")
        ch = num_channel
        self.dense_encoder = _dense_encoder(ch)
        for i in range(1, 5):
            blk = _Bag()
            blk.time_conformer = _conformer(ch)
            blk.freq_conformer = _conformer(ch)
            self.add_module(f"TSCB_{i}", blk)
        md = _Bag()
        md.dense_block = _dense_block(ch)
        md.sub_pixel = _sub_pixel(ch)
        md.conv_1 = nn.Conv2d(ch, 1, (1, 2))
        md.norm = nn.InstanceNorm2d(1, affine=True)
        md.prelu = nn.PReLU(1)
        md.final_conv = nn.Conv2d(1, 1, (1, 1))
        md.prelu_out = nn.PReLU(num_features, init=-0.25)
        self.mask_decoder = md
        cd = _Bag()
        cd.dense_block = _dense_block(ch)
        cd.sub_pixel = _sub_pixel(ch)
        cd.prelu = nn.PReLU(ch)
        cd.norm = nn.InstanceNorm2d(ch, affine=True)
        cd.conv = nn.Conv2d(ch, 2, (1, 2))
        self.complex_decoder = cd

        self.num_channel = ch
        self.num_features = num_features
        self.engine = ops.default_engine()     # "tcgen05" | "simt" main loop of the GEMM engine
        self.attention_variant = 0             # 0 tensor-core (mma.sync), 3 tcgen05 / TMEM kernel (attention_tc.cu), 1 SIMT cross-check
        # sequences at least this long take the tcgen05 kernel when attention_variant == 0 (measured, three-group kernel vs mma.sync:
        # 3.51 vs 4.91 ms at n = 4801, 7.15 vs 7.78 ms at n = 641, 3.6 vs 2.3 ms at n = 101 -- tools/attn_tc_check.py)
        self.attention_tc_min_len = 512
        self.half_v = True                     # tcgen05 engine: the depthwise output v is stored in fp16 between depthwise -> pw2 (-17 GB of HBM traffic per
                                               # 64 x 4 s step, pw2 7.3 -> 4.8 ms; whole-path cost 1.3-1.6e-4 of peak, profiles/r2/fp16_uv_error.txt)
        self.half_u = False                    # opt-in: the GLU output u in fp16 too -- measured a net loss (the depthwise kernel is FFMA-issue bound and
                                               # pays a half2 -> float2 conversion per window load: 13.0 vs 10.0 ms per step), so off by default
        self.overlap_decoders = False          # opt-in: complex decoder on a side stream next to the mask decoder (second buffer set; measured in DESIGN.md)
        self._packed: Optional[Dict[str, object]] = None
        self._packed_key = None
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}

    # -----------------------------------------------------------------------------------------
    # weight packing (once per parameter version / device)
    # -----------------------------------------------------------------------------------------
    def _version_key(self):
        dev = next(self.parameters()).device
        return (str(dev), tuple(int(p._version) for p in self.parameters()), tuple(int(b._version) for b in self.buffers()))

    def packed(self) -> Dict[str, object]:
        key = self._version_key()
        if self._packed is None or self._packed_key != key:
            self._packed = self._pack(next(self.parameters()).device)
            self._packed_key = key
        return self._packed

    def _pack(self, device) -> Dict[str, object]:
        sd = {k: v.detach().to("cpu", torch.float32) if v.is_floating_point() else v.detach().cpu() for k, v in self.state_dict().items()}
        P: Dict[str, object] = {}
        dev = lambda t: t.contiguous().to(device)

        def dense(prefix):
            for i in range(1, 5):
                P[f"{prefix}.conv{i}"] = pack_weight(conv_weight_matrix(sd[f"{prefix}.conv{i}.weight"]), 64, sd[f"{prefix}.conv{i}.bias"]).to(device)
                for nm in ("norm", "prelu"):
                    P[f"{prefix}.{nm}{i}.weight"] = dev(sd[f"{prefix}.{nm}{i}.weight"])
                P[f"{prefix}.norm{i}.bias"] = dev(sd[f"{prefix}.norm{i}.bias"])

        for e in self._encoder_names():
            P[f"{e}.conv_1.w"] = dev(sd[f"{e}.conv_1.0.weight"].reshape(64, 3))
            P[f"{e}.conv_1.b"] = dev(sd[f"{e}.conv_1.0.bias"])
            for k in ("conv_1.1.weight", "conv_1.1.bias", "conv_1.2.weight", "conv_2.1.weight", "conv_2.1.bias", "conv_2.2.weight"):
                P[f"{e}.{k}"] = dev(sd[f"{e}.{k}"])
            dense(f"{e}.dilated_dense")
            P[f"{e}.conv_2"] = pack_weight(conv_weight_matrix(sd[f"{e}.conv_2.0.weight"]), 64, sd[f"{e}.conv_2.0.bias"]).to(device)

        for i in range(1, 5):
            for ax in ("time", "freq"):
                p = f"TSCB_{i}.{ax}_conformer"
                for ff in ("ff1", "ff2"):
                    P[f"{p}.{ff}.w1"] = pack_weight(sd[f"{p}.{ff}.fn.fn.net.0.weight"], 64, sd[f"{p}.{ff}.fn.fn.net.0.bias"]).to(device)
                    P[f"{p}.{ff}.w2"] = pack_weight(sd[f"{p}.{ff}.fn.fn.net.3.weight"], 64, sd[f"{p}.{ff}.fn.fn.net.3.bias"]).to(device)
                    P[f"{p}.{ff}.ln"] = (dev(sd[f"{p}.{ff}.fn.norm.weight"]), dev(sd[f"{p}.{ff}.fn.norm.bias"]))
                wqkv = torch.cat([sd[f"{p}.attn.fn.to_q.weight"], sd[f"{p}.attn.fn.to_kv.weight"]], dim=0)
                P[f"{p}.attn.qkv"] = pack_weight(wqkv, 192, None).to(device)     # one wide n-tile: splitting N re-reads + re-normalises x (measured 2x slower)
                P[f"{p}.attn.out"] = pack_weight(sd[f"{p}.attn.fn.to_out.weight"], 64, sd[f"{p}.attn.fn.to_out.bias"]).to(device)
                P[f"{p}.attn.emb"] = dev(sd[f"{p}.attn.fn.rel_pos_emb.weight"])
                P[f"{p}.attn.emb_h"] = dev(ops.pack_rel_pos(sd[f"{p}.attn.fn.rel_pos_emb.weight"]))
                P[f"{p}.attn.ln"] = (dev(sd[f"{p}.attn.norm.weight"]), dev(sd[f"{p}.attn.norm.bias"]))
                P[f"{p}.conv.ln"] = (dev(sd[f"{p}.conv.net.0.weight"]), dev(sd[f"{p}.conv.net.0.bias"]))
                w1, b1 = glu_interleave(sd[f"{p}.conv.net.2.weight"].squeeze(-1), sd[f"{p}.conv.net.2.bias"])
                P[f"{p}.conv.pw1"] = pack_weight(w1, 256, b1).to(device)
                P[f"{p}.conv.dw"] = dev(sd[f"{p}.conv.net.4.conv.weight"].squeeze(1).t())          # [31][128] tap-major
                # BatchNorm1d(eval) folded with the depthwise bias: y = scale * conv + shift  (conformer.py:167)
                scale = sd[f"{p}.conv.net.5.weight"] / torch.sqrt(sd[f"{p}.conv.net.5.running_var"] + 1e-5)
                shift = sd[f"{p}.conv.net.5.bias"] + (sd[f"{p}.conv.net.4.conv.bias"] - sd[f"{p}.conv.net.5.running_mean"]) * scale
                P[f"{p}.conv.bn"] = (dev(scale), dev(shift))
                P[f"{p}.conv.pw2"] = pack_weight(sd[f"{p}.conv.net.7.weight"].squeeze(-1), 64, sd[f"{p}.conv.net.7.bias"]).to(device)
                P[f"{p}.post_norm"] = (dev(sd[f"{p}.post_norm.weight"]), dev(sd[f"{p}.post_norm.bias"]))

        for d in ("mask_decoder", "complex_decoder"):
            dense(f"{d}.dense_block")
            P[f"{d}.sub_pixel"] = pack_weight(conv_weight_matrix(sd[f"{d}.sub_pixel.conv.weight"]), 128, sd[f"{d}.sub_pixel.conv.bias"]).to(device)
        m = "mask_decoder"
        P[f"{m}.conv_1.w"] = dev(sd[f"{m}.conv_1.weight"][0, :, 0, :].t())                           # [2 taps][64]
        P[f"{m}.scalars"] = (float(sd[f"{m}.conv_1.bias"][0]), float(sd[f"{m}.norm.weight"][0]), float(sd[f"{m}.norm.bias"][0]),
                             float(sd[f"{m}.prelu.weight"][0]), float(sd[f"{m}.final_conv.weight"].reshape(-1)[0]),
                             float(sd[f"{m}.final_conv.bias"][0]))
        P[f"{m}.prelu_out"] = dev(sd[f"{m}.prelu_out.weight"])
        c = "complex_decoder"
        for k in ("prelu.weight", "norm.weight", "norm.bias", "conv.bias"):
            P[f"{c}.{k}"] = dev(sd[f"{c}.{k}"])
        P[f"{c}.conv.w"] = dev(sd[f"{c}.conv.weight"][:, :, 0, :].permute(0, 2, 1))                   # [2 out][2 taps][64]
        self._pack_extra(sd, P, dev, device)
        return P

    def _encoder_names(self):
        return ("dense_encoder",)

    def _pack_extra(self, sd, P, dev, device):
        """hook for variants with more parameters (tsc_diffusion.TSCNet: second encoder + MergeBlock)"""

    # -----------------------------------------------------------------------------------------
    # workspaces: one set of activation buffers per (device, B, T), reused across calls
    # -----------------------------------------------------------------------------------------
    def workspace(self, B: int, T: int, device) -> Dict[str, torch.Tensor]:
        key = (str(device), B, T)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        F, Fh = self.num_features, (self.num_features - 1) // 2 + 1
        P, Ph = B * T * F, B * T * Fh
        f32 = dict(device=device, dtype=torch.float32)
        ws = {
            # encoder @F=201: conv_1 output + 4 dense outputs + raw conv output
            "enc": [torch.empty(P, 64, **f32) for _ in range(5)], "enc_raw": torch.empty(P, 64, **f32),
            # decoders @F=101: 4 dense outputs + raw, sub-pixel output @F=202
            "dec": [torch.empty(Ph, 64, **f32) for _ in range(4)], "dec_raw": torch.empty(Ph, 64, **f32),
            "sp": torch.empty(B * T * 2 * Fh, 64, **f32),
            "xs": torch.empty(Ph, 64, **f32),          # TSCB output re-written in the pre-split conv-input format
            # conformer token buffers
            "x": torch.empty(Ph, 64, **f32), "y": torch.empty(Ph, 64, **f32), "h": torch.empty(Ph, 256, **f32),
            "qkv": torch.empty(Ph, 192, **f32), "o": torch.empty(Ph, 64, **f32),       # qkv doubles as the fp16 [Ph, 192] buffer
            "u": torch.empty(Ph, 128, **f32), "v": torch.empty(Ph, 128, **f32),
            # heads
            "mask_raw": torch.empty(B * T, F, **f32), "cplx": torch.empty(B * T, F, 2, **f32), "est": torch.empty(B * T, F, 2, **f32),
            "stats": torch.empty(B, 64, 2, **f32), "stats1": torch.empty(B, 1, 2, **f32),
            "in_ws": ops.inorm_workspace(B, T * 2 * Fh, 64, device),
        }
        if len(self._ws) > 4:          # keep the cache small: shapes change rarely in inference
            self._ws.clear()
        self._ws[key] = ws
        return ws

    # -----------------------------------------------------------------------------------------
    # building blocks
    # -----------------------------------------------------------------------------------------
    def _second_decoder_buffers(self, ws, B, T, device):
        if "dec2" not in ws:
            Fh = (self.num_features - 1) // 2 + 1
            Ph = B * T * Fh
            f32 = dict(device=device, dtype=torch.float32)
            ws["dec2"] = {"dec": [torch.empty(Ph, 64, **f32) for _ in range(4)], "dec_raw": torch.empty(Ph, 64, **f32),
                          "sp": torch.empty(B * T * 2 * Fh, 64, **f32), "stats": torch.empty(B, 64, 2, **f32),
                          "in_ws": ops.inorm_workspace(B, T * 2 * Fh, 64, device)}
        return ws["dec2"]

    def _side_stream(self, device):
        key = str(device)
        if not hasattr(self, "_side_streams"):
            self._side_streams = {}
        if key not in self._side_streams:
            self._side_streams[key] = torch.cuda.Stream(device=device)
        return self._side_streams[key]

    def _inorm_prelu(self, ws, raw, B, pix_per_b, gamma, beta, slope, out):
        ops.inorm_stats(raw, B, pix_per_b, 64, ws["stats"], ws["in_ws"])
        ops.inorm_prelu(raw, B, pix_per_b, ws["stats"], gamma, beta, slope, out)
        return out

    @staticmethod
    def _as_split(t: torch.Tensor) -> torch.Tensor:
        """fp32 [P, 64] storage viewed as the pre-split conv-input format: bfloat16 [P, 2, 64] (hi | lo), same bytes."""
        return t.view(torch.bfloat16).view(t.shape[0], 2, 64)

    def _conv_in(self, t: torch.Tensor) -> torch.Tensor:
        """the form in which conv inputs are stored / read for the active GEMM main loop"""
        return self._as_split(t) if self.engine == "tcgen05" else t

    def _dense(self, P, prefix, ws, x0, outs, raw, B, T, F):
        """DilatedDenseNet.forward (generator.py:24-32): layer i reads [out_{i-1}, ..., out_1, x0] through slot pointers.
        x0 / outs are conv-input tensors (pre-split bf16 on the tcgen05 engine, fp32 on the fp32 loop)."""
        loader = LOAD_CONV_SPLIT if self.engine == "tcgen05" else LOAD_CONV
        slots = [x0]
        for i in range(1, 5):
            w: PackedWeight = P[f"{prefix}.conv{i}"]
            ops.gemm(loader=loader, epilogue=EPI_BIAS, M=B * T * F, w=w, a=slots, out=raw, ldo=64, engine=self.engine, label="dconv",
                     conv=dict(B=B, T=T, Fin=F, Fout=F, taps_t=2, dil=2 ** (i - 1), stride_f=1, nslots=i))
            self._inorm_prelu(ws, raw, B, T * F, P[f"{prefix}.norm{i}.weight"], P[f"{prefix}.norm{i}.bias"], P[f"{prefix}.prelu{i}.weight"], outs[i - 1])
            slots = [outs[i - 1]] + slots
        return outs[3]

    def _conformer(self, P, p, ws, x, seq, M):
        """ConformerBlock.forward + the TSCB outer residual (conformer.py:206-212, generator.py:70,72); x updated in place."""
        eng = self.engine
        y, h, qkv, o, u, v = ws["y"], ws["h"], ws["qkv"], ws["o"], ws["u"], ws["v"]
        fused = eng == "tcgen05"
        # y = x + 0.5 * FF1(LN(x))
        if fused:
            ops.ffn_fused(x, y, P[f"{p}.ff1.ln"], P[f"{p}.ff1.w1"], P[f"{p}.ff1.w2"], 0.5)
        else:
            ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_SWISH, M=M, w=P[f"{p}.ff1.w1"], a=[x], lda=64, ln=P[f"{p}.ff1.ln"], out=h, ldo=256, engine=eng, label="ffn1")
            ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=P[f"{p}.ff1.w2"], a=[h], lda=256, out=y, ldo=64, resid=x, ldr=64, alpha=0.5, engine=eng, label="ffn2")
        # y += Attn(LN(y))
        if self.attention_variant in (0, 3):      # fp16 projection (q pre-scaled) feeding the tensor-core attention
            qkv_h = qkv.view(-1).view(torch.float16)[:M * 192].view(M, 192)
            ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_QKV_F16, M=M, w=P[f"{p}.attn.qkv"], a=[y], lda=64, ln=P[f"{p}.attn.ln"], out=qkv_h, ldo=192, engine=eng, label="qkv")
            variant = 3 if (self.attention_variant == 0 and seq.n >= self.attention_tc_min_len) else self.attention_variant
            ops.attention(qkv_h, P[f"{p}.attn.emb"], seq, o, variant, P[f"{p}.attn.emb_h"])
        else:                                # fp32 projection + fp32 SIMT attention (cross-check path)
            ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_BIAS, M=M, w=P[f"{p}.attn.qkv"], a=[y], lda=64, ln=P[f"{p}.attn.ln"], out=qkv, ldo=192, engine=eng, label="qkv")
            ops.attention(qkv, P[f"{p}.attn.emb"], seq, o, 1)
        ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=P[f"{p}.attn.out"], a=[o], lda=64, out=y, ldo=64, resid=y, ldr=64, alpha=1.0, engine=eng, label="attn_out")
        # y += ConvModule(y)
        hu, hv = fused and self.half_u, fused and self.half_v
        half = lambda t: t.view(-1).view(torch.float16)[:M * 128].view(M, 128)
        uu, vv = (half(u) if hu else u), (half(v) if hv else v)
        ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_GLU_F16 if hu else EPI_GLU, M=M, w=P[f"{p}.conv.pw1"], a=[y], lda=64, ln=P[f"{p}.conv.ln"], out=uu, ldo=128,
                 engine=eng, label="pw1_glu")
        ops.dwconv_bn_swish(uu, seq, P[f"{p}.conv.dw"], P[f"{p}.conv.bn"][0], P[f"{p}.conv.bn"][1], vv)
        ops.gemm(loader=LOAD_ROWS_F16 if hv else LOAD_ROWS, epilogue=EPI_RESID, M=M, w=P[f"{p}.conv.pw2"], a=[vv], lda=128, out=y, ldo=64, resid=y, ldr=64, alpha=1.0,
                 engine=eng, label="pw2")
        # x = post_norm(y + 0.5 * FF2(LN(y))) + x
        if fused:
            ops.ffn_fused(y, x, P[f"{p}.ff2.ln"], P[f"{p}.ff2.w1"], P[f"{p}.ff2.w2"], 0.5, post=P[f"{p}.post_norm"], resid2=x)
        else:
            ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_SWISH, M=M, w=P[f"{p}.ff2.w1"], a=[y], lda=64, ln=P[f"{p}.ff2.ln"], out=h, ldo=256, engine=eng, label="ffn1")
            ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=P[f"{p}.ff2.w2"], a=[h], lda=256, out=y, ldo=64, resid=y, ldr=64, alpha=0.5, engine=eng, label="ffn2")
            ops.layernorm_residual(y, P[f"{p}.post_norm"][0], P[f"{p}.post_norm"][1], x, x)
        return x

    # -----------------------------------------------------------------------------------------
    # forward
    # -----------------------------------------------------------------------------------------
    def forward_in3(self, in3: torch.Tensor, stages: Optional[dict] = None) -> torch.Tensor:
        """in3: [B, T, F, 3] = (|Y|, Re Y, Im Y) of the compressed spectrogram -> est [B*T, F, 2] (workspace tensor)."""
        if self.training:
            raise RuntimeError("se_b200.TSCNet.forward_in3 is the inference forward: call .eval() (train mode goes through forward())")
        B, T, F, _ = in3.shape
        if F != self.num_features or F % 2 == 0:
            raise RuntimeError(f"expected {self.num_features} frequency bins, got {F}")
        Fh = (F - 1) // 2 + 1
        dev = in3.device
        P = self.packed()
        ws = self.workspace(B, T, dev)
        x = self._encode(P, "dense_encoder", ws, in3, ws["x"])
        if stages is not None:
            stages["encoder"] = x.view(B, T, Fh, 64).clone()
        self._tscbs(P, ws, x, B, T, Fh, stages)
        return self._decode(P, ws, x, in3, stages)

    def _encode(self, P, e, ws, in3, x):
        """DenseEncoder.forward (generator.py:50-54): in3 [B, T, F, 3] -> x [B*T*F', 64] (written in place, returned)."""
        B, T, F, _ = in3.shape
        Fh = (F - 1) // 2 + 1
        eng = self.engine
        tc = eng == "tcgen05"
        conv_loader = LOAD_CONV_SPLIT if tc else LOAD_CONV
        enc, raw = [self._conv_in(t) for t in ws["enc"]], ws["enc_raw"]
        ops.conv1x1_in3(in3, P[f"{e}.conv_1.w"], P[f"{e}.conv_1.b"], raw)
        self._inorm_prelu(ws, raw, B, T * F, P[f"{e}.conv_1.1.weight"], P[f"{e}.conv_1.1.bias"], P[f"{e}.conv_1.2.weight"], enc[0])
        d4 = self._dense(P, f"{e}.dilated_dense", ws, enc[0], enc[1:5], raw, B, T, F)
        rawh = ws["dec_raw"]
        ops.gemm(loader=conv_loader, epilogue=EPI_BIAS, M=B * T * Fh, w=P[f"{e}.conv_2"], a=[d4], out=rawh, ldo=64, engine=eng, label="conv2",
                 conv=dict(B=B, T=T, Fin=F, Fout=Fh, taps_t=1, dil=1, stride_f=2, nslots=1))
        self._inorm_prelu(ws, rawh, B, T * Fh, P[f"{e}.conv_2.1.weight"], P[f"{e}.conv_2.1.bias"], P[f"{e}.conv_2.2.weight"], x)
        return x

    def _before_tscb(self, P, ws, x, i):
        """hook: tsc_diffusion.TSCNet runs its MergeBlock on x before every TSCB"""

    def _tscbs(self, P, ws, x, B, T, Fh, stages=None):
        """4 x TSCB (generator.py:67-74); x [B*T*F', 64] updated in place."""
        M = B * T * Fh
        seq_t = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh)
        seq_f = ops.make_seq(B * T, Fh, 1, Fh, 1)
        for i in range(1, 5):
            self._before_tscb(P, ws, x, i)
            self._conformer(P, f"TSCB_{i}.time_conformer", ws, x, seq_t, M)
            self._conformer(P, f"TSCB_{i}.freq_conformer", ws, x, seq_f, M)
            if stages is not None:
                stages[f"tscb{i}"] = x.view(B, T, Fh, 64).clone()

    def _decode(self, P, ws, x, in3, stages=None):
        """Mask / Complex decoders + recombination with the spectrogram in3 (generator.py:106-129,158-165) -> est [B*T, F, 2]."""
        B, T, F, _ = in3.shape
        Fh = (F - 1) // 2 + 1
        M = B * T * Fh
        dev = in3.device
        eng = self.engine
        tc = eng == "tcgen05"
        conv_loader = LOAD_CONV_SPLIT if tc else LOAD_CONV

        m, c = "mask_decoder", "complex_decoder"
        x0 = ops.split_planes(x, self._as_split(ws["xs"])) if tc else x      # both decoders read the TSCB output as a conv input
        b1, g_in, b_in, s1, wf, bf = P[f"{m}.scalars"]

        def mask_branch(w):        # MaskDecoder up to the InstanceNorm statistics (generator.py:106-109)
            dec, sp = [self._conv_in(t) for t in w["dec"]], w["sp"]
            d4 = self._dense(P, f"{m}.dense_block", w, x0, dec, w["dec_raw"], B, T, Fh)
            ops.gemm(loader=conv_loader, epilogue=EPI_SUBPIXEL, M=M, w=P[f"{m}.sub_pixel"], a=[d4], out=sp, ldo=64, engine=eng, label="subpixel",
                     conv=dict(B=B, T=T, Fin=Fh, Fout=Fh, taps_t=1, dil=1, stride_f=1, nslots=1))
            ops.mask_conv(sp, B * T, 2 * Fh, P[f"{m}.conv_1.w"], b1, ws["mask_raw"])
            ops.inorm_stats(ws["mask_raw"], B, T * F, 1, ws["stats1"], w["in_ws"])

        def complex_branch(w):     # ComplexDecoder (generator.py:124-129)
            dec, sp = [self._conv_in(t) for t in w["dec"]], w["sp"]
            d4 = self._dense(P, f"{c}.dense_block", w, x0, dec, w["dec_raw"], B, T, Fh)
            ops.gemm(loader=conv_loader, epilogue=EPI_SUBPIXEL, M=M, w=P[f"{c}.sub_pixel"], a=[d4], out=sp, ldo=64, engine=eng, label="subpixel",
                     conv=dict(B=B, T=T, Fin=Fh, Fout=Fh, taps_t=1, dil=1, stride_f=1, nslots=1))
            ops.inorm_stats(sp, B, T * 2 * Fh, 64, w["stats"], w["in_ws"])
            ops.complex_conv(sp, B, T, 2 * Fh, w["stats"], P[f"{c}.norm.weight"], P[f"{c}.norm.bias"], P[f"{c}.prelu.weight"],
                             P[f"{c}.conv.w"], P[f"{c}.conv.bias"], ws["cplx"])

        if self.overlap_decoders:
            # the two decoders are independent until the recombination: the complex decoder runs on a side stream with its own
            # buffers, so its bandwidth-bound norm kernels can fill the gaps of the other branch's tensor-bound convolutions
            w2 = self._second_decoder_buffers(ws, B, T, dev)
            main = torch.cuda.current_stream(dev)
            side = self._side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                complex_branch(w2)
            mask_branch(ws)
            main.wait_stream(side)
        else:
            mask_branch(ws)
            complex_branch(ws)

        # ---- mask tail + recombination (generator.py:110-112,158-165)
        mask_out = None
        if stages is not None:
            mask_out = torch.empty(B, T, F, device=dev, dtype=torch.float32)
            stages["complex"] = ws["cplx"].view(B, T, F, 2).clone()
        ops.mask_recombine(ws["mask_raw"], ws["stats1"], B, T, F, (g_in, b_in, s1, wf, bf), P[f"{m}.prelu_out"], in3, ws["cplx"],
                           ws["est"], mask_out)
        if stages is not None:
            stages["mask"] = mask_out
        return ws["est"]

    def _forward_train(self, x: torch.Tensor):
        """train-mode forward (core/function.py:221: dropout active, BatchNorm batch statistics + running update, autograd graph over the
        parameters): SURVEY 8f row f1, implemented in training.py (one autograd.Function around the generator, hand-written backward)"""
        from . import training
        return training.forward_train(self, x)

    def forward(self, x: torch.Tensor, diffusion_step=None):
        """x: complex64 (B, num_features, T) compressed spectrogram -> (final_real, final_imag), each fp32 (B, 1, T, F)."""
        if not x.is_cuda:
            raise RuntimeError("se_b200.TSCNet has no CPU path: input must be a CUDA tensor on an sm_100a device")
        if not x.is_complex():
            raise RuntimeError("TSCNet.forward expects the complex compressed spectrogram (B, F, T)")
        if self.training:
            return self._forward_train(x)
        with torch.no_grad(), torch.cuda.device(x.device):      # launches follow the input's device (cuda:k inputs)
            B, F, T = x.shape
            in3 = ops.spec_to_in3(x.to(torch.complex64))
            est = self.forward_in3(in3)
            fr = torch.empty(B, 1, T, F, device=x.device, dtype=torch.float32)
            fi = torch.empty(B, 1, T, F, device=x.device, dtype=torch.float32)
            ops.split_ri(est, fr, fi)
        return fr, fi
