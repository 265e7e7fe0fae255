"""GPU: every exported kernel against a plain PyTorch fp32 reference of the same op (or the oracle's function),
called through the C ABI.  Both main loops of the GEMM engine are exercised: `simt` (fp32 FFMA) and `tcgen05`
(split-bf16 tensor path); the tensor-core attention is checked against the oracle and against the SIMT variant."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, rel_max, rel_l2
from oracle import tscnet_oracle as O, weights

import se_b200
from se_b200 import ops, packing
from se_b200._lib import (EPI_BIAS, EPI_COMPRESS, EPI_GATE, EPI_GLU, EPI_QKV_F16, EPI_RESID, EPI_RESID_SCALE, EPI_SUBPIXEL, EPI_SWISH, LOAD_CONV,
                          LOAD_CONV_SPLIT, LOAD_HANKEL, LOAD_ROWS, LOAD_ROWS2, LOAD_ROWS_LN)

pytestmark = pytest.mark.gpu
DEV = "cuda"
ENGINES = ["simt", "tcgen05"]
TOL = {"simt": 2e-5, "tcgen05": 1e-4}


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


# ------------------------------------------------------------------------------------------------ GEMM engine
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("K,M", [(64, 300), (128, 128), (256, 1000)])
def test_gemm_rows_resid(engine, K, M):
    a, w, b, r = rnd(M, K, seed=1), rnd(64, K, seed=2, scale=K ** -0.5), rnd(64, seed=3), rnd(M, 64, seed=4)
    pw = packing.pack_weight(w.cpu(), 64, b.cpu()).to(DEV)
    out = torch.empty(M, 64, device=DEV)
    ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=pw, a=[a], lda=K, out=out, ldo=64, resid=r, ldr=64, alpha=0.5, engine=engine)
    ref = 0.5 * (a.double() @ w.double().t() + b.double()) + r.double()
    assert rel_max(out, ref) < TOL[engine]
    # in place on the residual buffer (how the conformer uses it)
    r2 = r.clone()
    ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=pw, a=[a], lda=K, out=r2, ldo=64, resid=r2, ldr=64, alpha=0.5, engine=engine)
    assert torch.equal(r2, out)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,groups", [(300, 1), (128 * 5, 5), (40007, 7)])
def test_gemm_merge_block(engine, M, groups):
    """MergeBlock's two contractions (models/tsc_diffusion.py:32-41): [x | cond] -> sigmoid * tanh gate with a per-group bias
    row (groups of rows_per_group consecutive rows; the last group may be ragged), then (x + W_o g + b_o) / sqrt(2)."""
    x, c = rnd(M, 64, seed=1), rnd(M, 64, seed=2)
    wm, wc, bm, bc = rnd(128, 64, seed=3, scale=0.125), rnd(128, 64, seed=4, scale=0.125), rnd(128, seed=5, scale=0.1), rnd(128, seed=6, scale=0.1)
    rpg = -(-M // groups)
    d = rnd(groups, 64, seed=7)
    wcat, bcat = packing.glu_interleave(torch.cat([wm, wc], 1).cpu(), (bm + bc).cpu())
    pw = packing.pack_weight(wcat, 128, bcat).to(DEV)
    rowbias = (d.double() @ wcat[:, :64].to(DEV).double().t()).float().contiguous()
    g = torch.empty(M, 64, device=DEV)
    ops.gemm(loader=LOAD_ROWS2, epilogue=EPI_GATE, M=M, w=pw, a=[x, c], lda=64, out=g, ldo=64, resid=rowbias, ldr=rpg, engine=engine)
    grp = torch.arange(M, device=DEV) // rpg
    y = (x.double() + d.double()[grp]) @ wm.double().t() + bm.double() + c.double() @ wc.double().t() + bc.double()
    ref = torch.sigmoid(y[:, :64]) * torch.tanh(y[:, 64:])
    assert rel_max(g, ref) < TOL[engine]
    # no row bias at all
    ops.gemm(loader=LOAD_ROWS2, epilogue=EPI_GATE, M=M, w=pw, a=[x, c], lda=64, out=g, ldo=64, engine=engine)
    y0 = x.double() @ wm.double().t() + bm.double() + c.double() @ wc.double().t() + bc.double()
    assert rel_max(g, torch.sigmoid(y0[:, :64]) * torch.tanh(y0[:, 64:])) < TOL[engine]
    # output residual, in place on x
    wo, bo = rnd(64, 64, seed=8, scale=0.125), rnd(64, seed=9, scale=0.1)
    po = packing.pack_weight(wo.cpu(), 64, bo.cpu()).to(DEV)
    gg = ref.float().contiguous()
    x2 = x.clone()
    ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID_SCALE, M=M, w=po, a=[gg], lda=64, out=x2, ldo=64, resid=x2, ldr=64, alpha=2 ** -0.5, engine=engine)
    ref2 = (x.double() + gg.double() @ wo.double().t() + bo.double()) / math.sqrt(2.0)
    assert rel_max(x2, ref2) < TOL[engine]


def test_diffusion_embed_matches_oracle():
    sd = weights.synth_state_dict(2, spec=weights.tsc_diffusion_spec())
    m = "merge_block"
    steps = torch.tensor([0.0, 7.0, 3.4, 48.75, 49.0])
    with torch.no_grad():
        e = O.diffusion_embedding(steps, sd, f"{m}.diffusion_embedding", 50)
        d_ref = F.linear(e, sd[f"{m}.diffusion_projection.weight"], sd[f"{m}.diffusion_projection.bias"])
        e_int = O.diffusion_embedding(torch.tensor([0, 7, 49]), sd, f"{m}.diffusion_embedding", 50)
        d_int = F.linear(e_int, sd[f"{m}.diffusion_projection.weight"], sd[f"{m}.diffusion_projection.bias"])
    wm = rnd(128, 64, seed=4, scale=0.125)
    dv = lambda k: sd[k].to(DEV).contiguous()
    d = torch.empty(5, 64, device=DEV)
    rb = torch.empty(5, 128, device=DEV)
    ops.diffusion_embed(steps.to(DEV), O.diffusion_step_table(50).to(DEV), dv(f"{m}.diffusion_embedding.projection1.weight"),
                        dv(f"{m}.diffusion_embedding.projection1.bias"), dv(f"{m}.diffusion_embedding.projection2.weight"),
                        dv(f"{m}.diffusion_embedding.projection2.bias"), dv(f"{m}.diffusion_projection.weight"), dv(f"{m}.diffusion_projection.bias"),
                        wm, d, rb)
    assert rel_max(d.cpu(), d_ref) < 1e-5
    assert rel_max(d.cpu()[[0, 1, 4]], d_int) < 1e-5
    assert rel_max(rb, d.double() @ wm.double().t()) < 1e-5


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_layernorm_loader_epilogues(engine):
    M = 777
    x = rnd(M, 64, seed=5, scale=3.0) + 0.7
    g, be = rnd(64, seed=6, scale=0.2) + 1.0, rnd(64, seed=7, scale=0.2)
    xn = F.layer_norm(x.double(), (64,), g.double(), be.double(), 1e-5)
    # Linear 64 -> 256 + Swish
    w, b = rnd(256, 64, seed=8, scale=0.17), rnd(256, seed=9, scale=0.1)
    out = torch.empty(M, 256, device=DEV)
    ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_SWISH, M=M, w=packing.pack_weight(w.cpu(), 256, b.cpu()).to(DEV), a=[x], lda=64, ln=(g, be), out=out, ldo=256, engine=engine)
    h = xn @ w.double().t() + b.double()
    assert rel_max(out, h * torch.sigmoid(h)) < TOL[engine]
    # pointwise conv 64 -> 256 + GLU (interleaved packing)
    wi, bi = packing.glu_interleave(w.cpu(), b.cpu())
    for ntile in (256, 64):           # 64: the 4-CTA-per-SM shape the model uses
        out = torch.empty(M, 128, device=DEV)
        ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_GLU, M=M, w=packing.pack_weight(wi, ntile, bi).to(DEV), a=[x], lda=64, ln=(g, be), out=out, ldo=128, engine=engine)
        assert rel_max(out, h[:, :128] * torch.sigmoid(h[:, 128:])) < TOL[engine]
    # q | k | v projection, no bias
    wq = rnd(192, 64, seed=10, scale=0.17)
    out = torch.empty(M, 192, device=DEV)
    ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_BIAS, M=M, w=packing.pack_weight(wq.cpu(), 192, None).to(DEV), a=[x], lda=64, ln=(g, be), out=out, ldo=192, engine=engine)
    assert rel_max(out, xn @ wq.double().t()) < TOL[engine]
    # the same projection in the fp16 layout the tensor-core attention reads (q pre-scaled)
    refh = xn @ wq.double().t()
    refh[:, :64] *= 0.25 * 1.4426950408889634
    for ntile in (192, 64):
        outh = torch.empty(M, 192, device=DEV, dtype=torch.float16)
        ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_QKV_F16, M=M, w=packing.pack_weight(wq.cpu(), ntile, None).to(DEV), a=[x], lda=64, ln=(g, be), out=outh, ldo=192, engine=engine)
        assert rel_max(outh.double(), refh) < 1e-3
    out = torch.empty(M, 192, device=DEV)
    ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_BIAS, M=M, w=packing.pack_weight(wq.cpu(), 64, None).to(DEV), a=[x], lda=64, ln=(g, be), out=out, ldo=192, engine=engine)
    assert rel_max(out, xn @ wq.double().t()) < TOL[engine]


@pytest.mark.parametrize("M", [128, 777, 40000])
def test_ffn_fused_tcgen05(M):
    """LayerNorm -> 64->256 -> Swish -> 256->64 -> *0.5 + x (-> post_norm + residual) in one tcgen05 kernel"""
    x = rnd(M, 64, seed=11, scale=2.0) + 0.3
    g, be = rnd(64, seed=12, scale=0.2) + 1.0, rnd(64, seed=13, scale=0.2)
    w1, b1 = rnd(256, 64, seed=14, scale=0.17), rnd(256, seed=15, scale=0.1)
    w2, b2 = rnd(64, 256, seed=16, scale=0.09), rnd(64, seed=17, scale=0.1)
    pg, pb, r2 = rnd(64, seed=18, scale=0.2) + 1.0, rnd(64, seed=19, scale=0.2), rnd(M, 64, seed=20)
    pw1 = packing.pack_weight(w1.cpu(), 64, b1.cpu()).to(DEV)
    pw2 = packing.pack_weight(w2.cpu(), 64, b2.cpu()).to(DEV)
    xd = x.double()
    h = F.layer_norm(xd, (64,), g.double(), be.double(), 1e-5) @ w1.double().t() + b1.double()
    y_ref = xd + 0.5 * ((h * torch.sigmoid(h)) @ w2.double().t() + b2.double())
    y = torch.empty_like(x)
    ops.ffn_fused(x, y, (g, be), pw1, pw2, 0.5)
    assert rel_max(y, y_ref) < TOL["tcgen05"]
    xi = x.clone()
    ops.ffn_fused(xi, xi, (g, be), pw1, pw2, 0.5)                     # in place
    assert torch.equal(xi, y)
    out_ref = F.layer_norm(y_ref, (64,), pg.double(), pb.double(), 1e-5) + r2.double()
    r2i = r2.clone()
    ops.ffn_fused(x, r2i, (g, be), pw1, pw2, 0.5, post=(pg, pb), resid2=r2i)   # out aliases resid2 (how the conformer calls it)
    assert rel_max(r2i, out_ref) < TOL["tcgen05"]


def _cl(x):      # (B, C, T, F) -> channels-last [B, T, F, C] contiguous
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("layer", [1, 2, 3, 4])
def test_gemm_dilated_conv(engine, layer):
    B, T, Fq = 2, 21, 13
    dil = 2 ** (layer - 1)
    slots_nchw = [rnd(B, 64, T, Fq, seed=20 + i) for i in range(layer)]            # newest first
    w = rnd(64, 64 * layer, 2, 3, seed=30, scale=(64 * layer * 6) ** -0.5)
    b = rnd(64, seed=31, scale=0.1)
    ref = F.conv2d(F.pad(torch.cat(slots_nchw, 1).double(), (1, 1, dil, 0)), w.double(), b.double(), dilation=(dil, 1))
    pw = packing.pack_weight(packing.conv_weight_matrix(w.cpu()), 64, b.cpu()).to(DEV)
    out = torch.empty(B * T * Fq, 64, device=DEV)
    ops.gemm(loader=LOAD_CONV, epilogue=EPI_BIAS, M=B * T * Fq, w=pw, a=[_cl(s) for s in slots_nchw], out=out, ldo=64, engine=engine,
             conv=dict(B=B, T=T, Fin=Fq, Fout=Fq, taps_t=2, dil=dil, stride_f=1, nslots=layer))
    assert rel_max(out.view(B, T, Fq, 64), _cl(ref)) < TOL[engine]


def _split(x_cl):      # fp32 [.., 64] -> pre-split conv-input format via the kernel
    flat = x_cl.reshape(-1, 64).contiguous()
    y = torch.empty(flat.shape[0], 2, 64, device=DEV, dtype=torch.bfloat16)
    ops.split_planes(flat, y)
    assert rel_max(y.float().sum(1), flat) < 2.0 ** -16
    return y


@pytest.mark.parametrize("layer,B,T,Fq", [(1, 2, 21, 13), (3, 1, 40, 101), (4, 2, 30, 201)])
def test_conv_on_presplit_activations_tcgen05(layer, B, T, Fq):
    """cp.async loader on (hi | lo) bf16 inputs: dilated dense conv, strided conv_2 and the sub-pixel conv"""
    dil = 2 ** (layer - 1)
    slots_nchw = [rnd(B, 64, T, Fq, seed=120 + i) for i in range(layer)]
    w = rnd(64, 64 * layer, 2, 3, seed=130, scale=(64 * layer * 6) ** -0.5)
    b = rnd(64, seed=131, scale=0.1)
    ref = F.conv2d(F.pad(torch.cat(slots_nchw, 1).double(), (1, 1, dil, 0)), w.double(), b.double(), dilation=(dil, 1))
    pw = packing.pack_weight(packing.conv_weight_matrix(w.cpu()), 64, b.cpu()).to(DEV)
    out = torch.empty(B * T * Fq, 64, device=DEV)
    ops.gemm(loader=LOAD_CONV_SPLIT, epilogue=EPI_BIAS, M=B * T * Fq, w=pw, a=[_split(_cl(s_)) for s_ in slots_nchw], out=out, ldo=64,
             engine="tcgen05", conv=dict(B=B, T=T, Fin=Fq, Fout=Fq, taps_t=2, dil=dil, stride_f=1, nslots=layer))
    assert rel_max(out.view(B, T, Fq, 64), _cl(ref)) < TOL["tcgen05"]
    if layer != 1:
        return
    Fh = (Fq - 1) // 2 + 1
    x = slots_nchw[0]
    w2, b2 = rnd(64, 64, 1, 3, seed=141, scale=0.07), rnd(64, seed=142, scale=0.1)
    ref = F.conv2d(x.double(), w2.double(), b2.double(), stride=(1, 2), padding=(0, 1))
    out = torch.empty(B * T * Fh, 64, device=DEV)
    ops.gemm(loader=LOAD_CONV_SPLIT, epilogue=EPI_BIAS, M=B * T * Fh, w=packing.pack_weight(packing.conv_weight_matrix(w2.cpu()), 64, b2.cpu()).to(DEV),
             a=[_split(_cl(x))], out=out, ldo=64, engine="tcgen05", conv=dict(B=B, T=T, Fin=Fq, Fout=Fh, taps_t=1, dil=1, stride_f=2, nslots=1))
    assert rel_max(out.view(B, T, Fh, 64), _cl(ref)) < TOL["tcgen05"]
    ws_, bs = rnd(128, 64, 1, 3, seed=144, scale=0.07), rnd(128, seed=145, scale=0.1)
    ref = O.sub_pixel(x.cpu(), {"p.conv.weight": ws_.cpu(), "p.conv.bias": bs.cpu()}, "p")
    out = torch.empty(B * T * 2 * Fq, 64, device=DEV)
    ops.gemm(loader=LOAD_CONV_SPLIT, epilogue=EPI_SUBPIXEL, M=B * T * Fq, w=packing.pack_weight(packing.conv_weight_matrix(ws_.cpu()), 128, bs.cpu()).to(DEV),
             a=[_split(_cl(x))], out=out, ldo=64, engine="tcgen05", conv=dict(B=B, T=T, Fin=Fq, Fout=Fq, taps_t=1, dil=1, stride_f=1, nslots=1))
    assert rel_max(out.view(B, T, 2 * Fq, 64).cpu(), _cl(ref)) < TOL["tcgen05"]


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_strided_conv_and_subpixel(engine):
    B, T, Fq = 2, 9, 21
    Fh = (Fq - 1) // 2 + 1
    x = rnd(B, 64, T, Fq, seed=40)
    w, b = rnd(64, 64, 1, 3, seed=41, scale=0.07), rnd(64, seed=42, scale=0.1)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=(1, 2), padding=(0, 1))
    out = torch.empty(B * T * Fh, 64, device=DEV)
    ops.gemm(loader=LOAD_CONV, epilogue=EPI_BIAS, M=B * T * Fh, w=packing.pack_weight(packing.conv_weight_matrix(w.cpu()), 64, b.cpu()).to(DEV),
             a=[_cl(x)], out=out, ldo=64, engine=engine, conv=dict(B=B, T=T, Fin=Fq, Fout=Fh, taps_t=1, dil=1, stride_f=2, nslots=1))
    assert rel_max(out.view(B, T, Fh, 64), _cl(ref)) < TOL[engine]
    # sub-pixel (generator.py:85-92)
    xs = rnd(B, 64, T, Fh, seed=43)
    ws, bs = rnd(128, 64, 1, 3, seed=44, scale=0.07), rnd(128, seed=45, scale=0.1)
    sd = {"p.conv.weight": ws.cpu(), "p.conv.bias": bs.cpu()}
    ref = O.sub_pixel(xs.cpu(), sd, "p")
    out = torch.empty(B * T * 2 * Fh, 64, device=DEV)
    ops.gemm(loader=LOAD_CONV, epilogue=EPI_SUBPIXEL, M=B * T * Fh, w=packing.pack_weight(packing.conv_weight_matrix(ws.cpu()), 128, bs.cpu()).to(DEV),
             a=[_cl(xs)], out=out, ldo=64, engine=engine, conv=dict(B=B, T=T, Fin=Fh, Fout=Fh, taps_t=1, dil=1, stride_f=1, nslots=1))
    assert rel_max(out.view(B, T, 2 * Fh, 64).cpu(), _cl(ref)) < TOL[engine]


# ------------------------------------------------------------------------------------------------ DSP bracket
@pytest.mark.parametrize("engine", ENGINES)
def test_compressed_stft_matches_oracle(engine):
    x, _ = weights.synth_wave(3, 4800, seed=3, kind="speech")
    ref = O.compressed_stft(x)
    got = se_b200.compressed_stft(x.to(DEV), 400, 100, torch.hamming_window(400).to(DEV), engine=engine).cpu()
    assert got.shape == ref.shape and got.dtype == torch.complex64
    # BASELINE tolerance: compressed spectrogram within 1e-4 relative (rel-L2; worst-bin/peak reported alongside)
    assert rel_l2(torch.view_as_real(got), torch.view_as_real(ref)) < 1e-4
    # worst bin / peak: both engines are fp32-grade (the tensor path uses three bf16 planes per operand, six products)
    assert rel_max(torch.view_as_real(got), torch.view_as_real(ref)) < 2e-4


@pytest.mark.parametrize("engine", ENGINES)
def test_uncompressed_istft_matches_oracle(engine):
    g = torch.Generator().manual_seed(4)
    spec = torch.complex(torch.randn(2, 201, 33, generator=g), torch.randn(2, 201, 33, generator=g))
    ref = O.uncompressed_istft(spec)
    got = se_b200.uncompressed_istft(spec.to(DEV), 400, 100, torch.hamming_window(400).to(DEV), engine=engine).cpu()
    assert got.shape == ref.shape
    assert rel_max(got, ref) < 2e-5


def test_rms_pad_and_layout_kernels():
    x, _ = weights.synth_wave(3, 1950, seed=5, kind="noise")
    xpad, c = ops.rms_pad(x.to(DEV), 2000, normalize=True)
    cr = torch.sqrt(1950 / torch.sum(x ** 2.0, dim=-1))
    assert rel_max(c.cpu(), cr) < 1e-6
    xs = x * cr[:, None]
    xs = torch.cat([xs, xs[:, :50]], -1)
    ref = F.pad(xs.unsqueeze(1), (200, 200), mode="reflect").squeeze(1)
    assert rel_max(xpad.cpu(), ref) < 1e-6
    g = torch.Generator().manual_seed(6)
    spec = torch.complex(torch.randn(2, 201, 37, generator=g), torch.randn(2, 201, 37, generator=g)).to(DEV)
    in3 = ops.spec_to_in3(spec)
    assert torch.equal(in3[..., 1], spec.real.permute(0, 2, 1)) and torch.equal(in3[..., 2], spec.imag.permute(0, 2, 1))
    assert rel_max(in3[..., 0], spec.abs().permute(0, 2, 1)) < 1e-6
    assert torch.equal(ops.in3_to_spec(in3), spec)


# ------------------------------------------------------------------------------------------------ bandwidth kernels
def test_conv1x1_inorm_prelu():
    B, T, Fq = 3, 50, 201
    in3 = rnd(B, T, Fq, 3, seed=50)
    w, b = rnd(64, 3, seed=51), rnd(64, seed=52)
    raw = torch.empty(B * T * Fq, 64, device=DEV)
    ops.conv1x1_in3(in3, w, b, raw)
    ref = in3.double() @ w.double().t() + b.double()
    assert rel_max(raw.view(B, T, Fq, 64), ref) < 1e-6
    g, be, sl = rnd(64, seed=53) * 0.1 + 1, rnd(64, seed=54) * 0.1, rnd(64, seed=55) * 0.05 + 0.25
    x = raw.view(B, T * Fq, 64) * 2.0 + 5.0            # a mean well away from zero stresses the variance
    x = x.contiguous()
    stats = torch.empty(B, 64, 2, device=DEV)
    wsb = ops.inorm_workspace(B, T * Fq, 64, DEV)
    ops.inorm_stats(x, B, T * Fq, 64, stats, wsb)
    y = torch.empty_like(x)
    ops.inorm_prelu(x, B, T * Fq, stats, g, be, sl, y)
    xn = x.permute(0, 2, 1).reshape(B, 64, T, Fq)
    ref = F.prelu(F.instance_norm(xn.double(), weight=g.double(), bias=be.double(), eps=1e-5), sl.double())
    assert rel_max(y.view(B, T, Fq, 64), ref.permute(0, 2, 3, 1)) < 1e-5
    ys = torch.empty(B * T * Fq, 2, 64, device=DEV, dtype=torch.bfloat16)          # pre-split conv-input format
    ops.inorm_prelu(x, B, T * Fq, stats, g, be, sl, ys)
    assert rel_max(ys.float().sum(1).view(B, T * Fq, 64), y) < 2.0 ** -16
    stats2 = torch.empty_like(stats)
    ops.inorm_stats(x, B, T * Fq, 64, stats2, wsb)
    assert torch.equal(stats, stats2)                  # deterministic reduction


def test_heads_and_recombine():
    B, T, Fq = 2, 17, 201
    Fh = 101
    sd = weights.synth_state_dict(2)
    sp = rnd(B, 64, T, 2 * Fh, seed=60)
    in3 = rnd(B, T, Fq, 3, seed=61)
    spc = _cl(sp)
    # mask head
    m = "mask_decoder"
    mraw = torch.empty(B * T, Fq, device=DEV)
    ops.mask_conv(spc, B * T, 2 * Fh, sd[f"{m}.conv_1.weight"][0, :, 0, :].t().contiguous().to(DEV), float(sd[f"{m}.conv_1.bias"]), mraw)
    ref_raw = F.conv2d(sp.cpu().double(), sd[f"{m}.conv_1.weight"].double(), sd[f"{m}.conv_1.bias"].double())
    assert rel_max(mraw.view(B, T, Fq).cpu(), ref_raw[:, 0]) < 1e-5
    st1 = torch.empty(B, 1, 2, device=DEV)
    wsb = ops.inorm_workspace(B, T * 2 * Fh, 64, DEV)
    ops.inorm_stats(mraw, B, T * Fq, 1, st1, wsb)
    # complex head
    c = "complex_decoder"
    st = torch.empty(B, 64, 2, device=DEV)
    ops.inorm_stats(spc, B, T * 2 * Fh, 64, st, wsb)
    cplx = torch.empty(B * T, Fq, 2, device=DEV)
    ops.complex_conv(spc, B, T, 2 * Fh, st, sd[f"{c}.norm.weight"].to(DEV), sd[f"{c}.norm.bias"].to(DEV), sd[f"{c}.prelu.weight"].to(DEV),
                     sd[f"{c}.conv.weight"][:, :, 0, :].permute(0, 2, 1).contiguous().to(DEV), sd[f"{c}.conv.bias"].to(DEV), cplx)
    h = F.prelu(F.instance_norm(sp.cpu(), weight=sd[f"{c}.norm.weight"], bias=sd[f"{c}.norm.bias"], eps=1e-5), sd[f"{c}.prelu.weight"])
    ref_c = F.conv2d(h, sd[f"{c}.conv.weight"], sd[f"{c}.conv.bias"])                 # (B, 2, T, F)
    assert rel_max(cplx.view(B, T, Fq, 2).cpu(), ref_c.permute(0, 2, 3, 1)) < 2e-5
    # recombine
    est = torch.empty(B * T, Fq, 2, device=DEV)
    mask = torch.empty(B, T, Fq, device=DEV)
    scal = (float(sd[f"{m}.norm.weight"]), float(sd[f"{m}.norm.bias"]), float(sd[f"{m}.prelu.weight"]),
            float(sd[f"{m}.final_conv.weight"]), float(sd[f"{m}.final_conv.bias"]))
    ops.mask_recombine(mraw, st1, B, T, Fq, scal, sd[f"{m}.prelu_out.weight"].to(DEV), in3, cplx, est, mask)
    hm = F.prelu(F.instance_norm(ref_raw.float(), weight=sd[f"{m}.norm.weight"], bias=sd[f"{m}.norm.bias"], eps=1e-5), sd[f"{m}.prelu.weight"])
    hm = F.conv2d(hm, sd[f"{m}.final_conv.weight"], sd[f"{m}.final_conv.bias"]).permute(0, 3, 2, 1).squeeze(-1)
    ref_mask = F.prelu(hm, sd[f"{m}.prelu_out.weight"]).permute(0, 2, 1)                 # (B, T, F)
    assert rel_max(mask.cpu(), ref_mask) < 2e-5
    ref_est = ref_mask.unsqueeze(-1) * in3.cpu()[..., 1:3] + ref_c.permute(0, 2, 3, 1)
    assert rel_max(est.view(B, T, Fq, 2).cpu(), ref_est) < 2e-5
    re, im = torch.empty(B, 1, T, Fq, device=DEV), torch.empty(B, 1, T, Fq, device=DEV)
    ops.split_ri(est, re, im)
    assert torch.equal(re.view(B, T, Fq), est.view(B, T, Fq, 2)[..., 0]) and torch.equal(im.view(B, T, Fq), est.view(B, T, Fq, 2)[..., 1])


def test_layernorm_residual():
    x, r = rnd(1000, 64, seed=70, scale=2.0) + 1.0, rnd(1000, 64, seed=71)
    g, b = rnd(64, seed=72) * 0.1 + 1, rnd(64, seed=73) * 0.1
    out = torch.empty_like(x)
    ops.layernorm_residual(x, g, b, r, out)
    assert rel_max(out, F.layer_norm(x.double(), (64,), g.double(), b.double(), 1e-5) + r.double()) < 1e-5
    r2 = r.clone()
    ops.layernorm_residual(x, g, b, r2, r2)
    assert torch.equal(r2, out)


# ------------------------------------------------------------------------------------------------ sequence kernels
def _seq_layouts(B, T, Fh):
    return {"time": (ops.make_seq(B * Fh, T, Fh, T * Fh, Fh), lambda x: x.permute(0, 2, 1, 3).reshape(B * Fh, T, -1),
                     lambda y: y.reshape(B, Fh, T, -1).permute(0, 2, 1, 3)),
            "freq": (ops.make_seq(B * T, Fh, 1, Fh, 1), lambda x: x.reshape(B * T, Fh, -1), lambda y: y.reshape(B, T, Fh, -1))}


def _attention_core_ref(qkv_seq, emb):
    S, n, _ = qkv_seq.shape
    q, k, v = (qkv_seq[..., i * 64:(i + 1) * 64].reshape(S, n, 4, 16).permute(0, 2, 1, 3).double() for i in range(3))
    pos = torch.arange(n)
    dist = (pos[:, None] - pos[None, :]).clamp(-512, 512) + 512
    dots = (q @ k.transpose(-1, -2) + torch.einsum("bhnd,nrd->bhnr", q, emb.double()[dist])) * 0.25
    return (dots.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(S, n, 64)


@pytest.mark.parametrize("variant", [1, 0, 3])     # 3 = the tcgen05 / TMEM kernel (attention_tc.cu)
@pytest.mark.parametrize("axis,B,T,Fh", [("freq", 2, 5, 101), ("time", 1, 150, 3), ("time", 1, 700, 2), ("freq", 1, 3, 64),
                                          ("time", 1, 641, 2), ("time", 1, 129, 2), ("time", 1, 128, 2), ("time", 1, 1409, 1), ("time", 1, 65, 1), ("time", 2, 97, 1), ("time", 1, 3, 2), ("freq", 1, 2, 9), ("time", 1, 40, 1), ("time", 1, 1400, 1)])     # 641 = 10 x 64 + 1, 129, 65, 1409 (far-field tiles): the last key enters the tcgen05 kernel's initial softmax state; 16-key tail body, 3-warp CTAs (mma.sync kernel); 1400 > 2*512 + 128: far-field (clamped) tiles take the constant shortcut
def test_attention(variant, axis, B, T, Fh):
    qkv = rnd(B, T, Fh, 192, seed=80, scale=1.5)
    emb = rnd(1025, 16, seed=81)
    seq, to_seq, from_seq = _seq_layouts(B, T, Fh)[axis]
    out = torch.zeros(B * T * Fh, 64, device=DEV)
    inp = qkv.view(-1, 192)
    if variant != 1:      # what SEB_EPI_QKV_F16 writes: fp16, q pre-scaled by dim_head^-0.5 * log2(e)
        inp = torch.cat([inp[:, :64] * (0.25 * 1.4426950408889634), inp[:, 64:]], 1).to(torch.float16).contiguous()
    ops.attention(inp, emb, seq, out, variant)
    ref = from_seq(_attention_core_ref(to_seq(qkv.cpu()), emb.cpu()))
    tol = 1e-5 if variant == 1 else 2e-3
    assert rel_max(out.view(B, T, Fh, 64).cpu(), ref) < tol


@pytest.mark.parametrize("variant", [0, 3])
def test_attention_peaky_logits(variant):
    """large, sharply peaked logits: the lazy running maximum of variant 0 has to move its reference (and rescale) often"""
    B, T, Fh = 1, 300, 2
    qkv = rnd(B, T, Fh, 192, seed=82, scale=4.0)
    emb = rnd(1025, 16, seed=83, scale=2.0)
    seq, to_seq, from_seq = _seq_layouts(B, T, Fh)["time"]
    out = torch.zeros(B * T * Fh, 64, device=DEV)
    inp = qkv.view(-1, 192)
    inp = torch.cat([inp[:, :64] * (0.25 * 1.4426950408889634), inp[:, 64:]], 1).to(torch.float16).contiguous()
    ops.attention(inp, emb, seq, out, variant)
    assert torch.isfinite(out).all()
    # reference on the SAME fp16-rounded operands (the logits are too steep for the operand rounding itself to be negligible)
    qh = inp.float()
    qkv_r = torch.cat([qh[:, :64] / (0.25 * 1.4426950408889634), qh[:, 64:]], 1).view(B, T, Fh, 192)
    ref = from_seq(_attention_core_ref(to_seq(qkv_r.cpu()), emb.to(torch.float16).float().cpu()))
    assert rel_max(out.view(B, T, Fh, 64).cpu(), ref) < 5e-3


@pytest.mark.parametrize("axis,B,T,Fh", [("freq", 2, 5, 101), ("time", 2, 130, 3)])
def test_dwconv_bn_swish(axis, B, T, Fh):
    sd = weights.synth_state_dict(1)
    p = "TSCB_1.time_conformer.conv.net"
    u = rnd(B, T, Fh, 128, seed=90)
    seq, to_seq, from_seq = _seq_layouts(B, T, Fh)[axis]
    scale = sd[f"{p}.5.weight"] / torch.sqrt(sd[f"{p}.5.running_var"] + 1e-5)
    shift = sd[f"{p}.5.bias"] + (sd[f"{p}.4.conv.bias"] - sd[f"{p}.5.running_mean"]) * scale
    y = torch.empty_like(u)
    ops.dwconv_bn_swish(u.view(-1, 128), seq, sd[f"{p}.4.conv.weight"].squeeze(1).t().contiguous().to(DEV), scale.to(DEV), shift.to(DEV), y.view(-1, 128))
    h = to_seq(u.cpu()).transpose(1, 2)
    h = F.conv1d(F.pad(h, (15, 15)), sd[f"{p}.4.conv.weight"], sd[f"{p}.4.conv.bias"], groups=128)
    h = F.batch_norm(h, sd[f"{p}.5.running_mean"], sd[f"{p}.5.running_var"], sd[f"{p}.5.weight"], sd[f"{p}.5.bias"], False, 0.0, 1e-5)
    ref = from_seq((h * torch.sigmoid(h)).transpose(1, 2))
    assert rel_max(y.cpu(), ref) < 1e-5




# ---- fp16 storage of the conv module's intermediates (GLU output u, depthwise output v): half the HBM traffic of pw1 -> depthwise -> pw2
@pytest.mark.parametrize("engine", ["simt", "tcgen05"])
@pytest.mark.parametrize("M", [129, 777, 20000])
def test_glu_fp16_output_and_fp16_rows_gemm(engine, M):
    from se_b200._lib import EPI_GLU_F16, LOAD_ROWS_F16
    x = rnd(M, 64, seed=205, scale=3.0) + 0.7
    g, be = rnd(64, seed=206, scale=0.2) + 1.0, rnd(64, seed=207, scale=0.2)
    xn = F.layer_norm(x.double(), (64,), g.double(), be.double(), 1e-5)
    w, b = rnd(256, 64, seed=208, scale=0.17), rnd(256, seed=209, scale=0.1)
    h = xn @ w.double().t() + b.double()
    wi, bi = packing.glu_interleave(w.cpu(), b.cpu())
    u = torch.empty(M, 128, device=DEV, dtype=torch.float16)
    ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_GLU_F16, M=M, w=packing.pack_weight(wi, 256, bi).to(DEV), a=[x], lda=64, ln=(g, be), out=u, ldo=128, engine=engine)
    ref_u = h[:, :128] * torch.sigmoid(h[:, 128:])
    assert rel_max(u.double(), ref_u) < 1e-3                     # fp16 rounding of the output: 2^-11 relative
    # pointwise 128 -> 64 + bias + residual reading fp16 rows (exact in the bf16 hi | lo split)
    v = (rnd(M, 128, seed=210, scale=1.5)).to(torch.float16)
    w3, b3, res = rnd(64, 128, seed=211, scale=0.12), rnd(64, seed=212, scale=0.1), rnd(M, 64, seed=213)
    out = res.clone()
    ops.gemm(loader=LOAD_ROWS_F16, epilogue=EPI_RESID, M=M, w=packing.pack_weight(w3.cpu(), 64, b3.cpu()).to(DEV), a=[v], lda=128, out=out, ldo=64, resid=out, ldr=64,
             alpha=1.0, engine=engine)
    ref = v.double() @ w3.double().t() + b3.double() + res.double()
    assert rel_max(out, ref) < TOL[engine]


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("axis,B,T,Fh", [("freq", 2, 5, 101), ("time", 2, 130, 3), ("time", 1, 641, 2)])
def test_dwconv_bn_swish_fp16_io(axis, B, T, Fh, in_dtype):
    sd = weights.synth_state_dict(1)
    p = "TSCB_1.time_conformer.conv.net"
    u = rnd(B, T, Fh, 128, seed=290).to(in_dtype)
    seq, to_seq, from_seq = _seq_layouts(B, T, Fh)[axis]
    scale = sd[f"{p}.5.weight"] / torch.sqrt(sd[f"{p}.5.running_var"] + 1e-5)
    shift = sd[f"{p}.5.bias"] + (sd[f"{p}.4.conv.bias"] - sd[f"{p}.5.running_mean"]) * scale
    y = torch.empty(u.shape, device=DEV, dtype=torch.float16)
    ops.dwconv_bn_swish(u.view(-1, 128), seq, sd[f"{p}.4.conv.weight"].squeeze(1).t().contiguous().to(DEV), scale.to(DEV), shift.to(DEV), y.view(-1, 128))
    h = to_seq(u.float().cpu()).transpose(1, 2)
    h = F.conv1d(F.pad(h, (15, 15)), sd[f"{p}.4.conv.weight"], sd[f"{p}.4.conv.bias"], groups=128)
    h = F.batch_norm(h, sd[f"{p}.5.running_mean"], sd[f"{p}.5.running_var"], sd[f"{p}.5.weight"], sd[f"{p}.5.bias"], False, 0.0, 1e-5)
    ref = from_seq((h * torch.sigmoid(h)).transpose(1, 2))
    assert rel_max(y.float().cpu(), ref) < 1e-3


def test_token_gemm_tma_and_cp_async_staging_are_bit_identical(tmp_path):
    """the token GEMMs stage contiguous 256-byte rows by tensor-map TMA (default) or by cp.async (SEB200_TOK_NO_TMA=1, read once per process): same
    arithmetic on the same values, so the two paths must agree bit for bit -- including the hardware zero fill of the rows past M (M = 777, 129)"""
    import subprocess
    import sys
    script = r'''
import sys, torch
sys.path.insert(0, %r)
import se_b200
from se_b200 import ops, packing
from se_b200._lib import EPI_GLU, EPI_QKV_F16, EPI_RESID, LOAD_ROWS, LOAD_ROWS_LN, LOAD_ROWS_F16
torch.manual_seed(3)
dev = "cuda"
outs = []
for M in (777, 129, 4096):
    x = torch.randn(M, 64, device=dev) * 2 + 0.3
    g, be = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
    wq = packing.pack_weight(torch.randn(192, 64) * 0.17, 192, None).to(dev)
    o1 = torch.empty(M, 192, device=dev, dtype=torch.float16)
    ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_QKV_F16, M=M, w=wq, a=[x], lda=64, ln=(g, be), out=o1, ldo=192, engine="tcgen05")
    wi, bi = packing.glu_interleave(torch.randn(256, 64) * 0.17, torch.randn(256) * 0.1)
    o2 = torch.empty(M, 128, device=dev)
    ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_GLU, M=M, w=packing.pack_weight(wi, 256, bi).to(dev), a=[x], lda=64, ln=(g, be), out=o2, ldo=128, engine="tcgen05")
    o3 = torch.randn(M, 64, device=dev)
    ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=packing.pack_weight(torch.randn(64, 64) * 0.17, 64, torch.randn(64) * 0.1).to(dev), a=[x], lda=64,
             resid=o3, ldr=64, out=o3, ldo=64, engine="tcgen05")
    v = torch.randn(M, 128, device=dev).to(torch.float16)
    o4 = torch.randn(M, 64, device=dev)
    ops.gemm(loader=LOAD_ROWS_F16, epilogue=EPI_RESID, M=M, w=packing.pack_weight(torch.randn(64, 128) * 0.12, 64, torch.randn(64) * 0.1).to(dev), a=[v], lda=128,
             resid=o4, ldr=64, out=o4, ldo=64, engine="tcgen05")
    outs += [o1.float().cpu(), o2.cpu(), o3.cpu(), o4.cpu()]
torch.save(outs, sys.argv[1])
''' % ROOT
    files = []
    for flag in ("0", "1"):
        f = str(tmp_path / f"tok_{flag}.pt")
        env = dict(os.environ, SEB200_TOK_NO_TMA=flag)
        r = subprocess.run([sys.executable, "-c", script, f], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        files.append(torch.load(f))
    assert len(files[0]) == 12
    for a, b in zip(*files):
        assert torch.isfinite(a).all() and torch.equal(a, b)
