"""CPU oracle for the SCP-GAN / CMGAN generator hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement (torch on CPU, no nn.Module classes,
no einops) of the reference's algorithm for the path named in BASELINE.json:

    predict()              /root/reference/inference_gan.py:75-100
    compressed_stft()      /root/reference/core/function.py:685-693
    power_compress()       /root/reference/core/function.py:625-634
    TSCNet.forward()       /root/reference/models/generator.py:145-167
    ConformerBlock.forward /root/reference/models/conformer.py:206-212
    power_uncompress()     /root/reference/core/function.py:636-645
    uncompressed_istft()   /root/reference/core/function.py:695-703
    normalize_batch() / batch_stft()  /root/reference/core/function.py:647-683
    tsc_diffusion.TSCNet / MergeBlock /root/reference/models/tsc_diffusion.py:16-90 (+ models/DiffuSE.py:39-69), SURVEY 8f row f3

It is driven purely by a ``state_dict`` with the reference's 359 keys.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import it; the product package never does.

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/make_golden.py`` imports the unmodified reference from
/root/reference (with the five third-party stubs of ``oracle/ref_import.py``),
runs ``inference_gan.predict`` and ``TSCNet`` on seeded inputs and commits the
results under ``tests/golden/``; ``tests/test_oracle.py`` checks this
restatement against those files (and, when /root/reference is mounted, against
the live reference, stage by stage).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

N_FFT = 400          # config/default.py:19-23
HOP = 100
N_BINS = N_FFT // 2 + 1
COMPRESS_EXP = 0.3   # core/function.py:628-629
MAX_POS = 512        # models/conformer.py:81
HEADS = 4            # models/generator.py:60-65
DIM_HEAD = 16

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# DSP bracket
# ----------------------------------------------------------------------------
def hamming_periodic(n: int = N_FFT) -> torch.Tensor:
    """torch.hamming_window(n) (periodic): inference_gan.py:78."""
    k = torch.arange(n, dtype=torch.float64)
    return (0.54 - 0.46 * torch.cos(2.0 * math.pi * k / n)).to(torch.float32)


def stft_frames(x: torch.Tensor, n_fft: int = N_FFT, hop: int = HOP) -> torch.Tensor:
    """Centre (reflect) padding + framing as torch.stft does it by default
    (core/function.py:690-691).  x: (B, L) -> (B, T, n_fft), T = L // hop + 1."""
    pad = n_fft // 2
    xp = F.pad(x.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    return xp.unfold(-1, n_fft, hop)


def power_compress(spec: torch.Tensor) -> torch.Tensor:
    """|S|^0.3 * exp(j*angle(S))  (core/function.py:625-634), written the way
    the reference evaluates it: abs, angle, pow, cos, sin."""
    mag = spec.abs() ** COMPRESS_EXP
    ph = spec.angle()
    return torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))


def power_uncompress(spec: torch.Tensor) -> torch.Tensor:
    """|S|^(1/0.3) * exp(j*angle(S))  (core/function.py:636-645)."""
    mag = spec.abs() ** (1.0 / COMPRESS_EXP)
    ph = spec.angle()
    return torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))


def compressed_stft(x: torch.Tensor, n_fft: int = N_FFT, hop: int = HOP,
                    window: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, L) fp32 -> complex64 (B, n_fft/2+1, T).  core/function.py:685-693."""
    w = hamming_periodic(n_fft).to(x.device) if window is None else window
    fr = stft_frames(x, n_fft, hop) * w
    spec = torch.fft.rfft(fr, dim=-1).transpose(1, 2)
    return power_compress(spec)


def normalize_batch(clean: torch.Tensor, noisy: torch.Tensor):
    """core/function.py:647-659: c = sqrt(L / sum noisy^2) per utterance, applied to both signals."""
    c = torch.sqrt(noisy.size(-1) / torch.sum(noisy ** 2.0, dim=-1, keepdim=True))
    return clean * c, noisy * c


def batch_stft(clean: torch.Tensor, noisy: torch.Tensor):
    """core/function.py:664-683 (forward DSP of the training caller): normalised waveforms and their compressed STFTs."""
    clean, noisy = normalize_batch(clean, noisy)
    return clean, noisy, compressed_stft(clean), compressed_stft(noisy)


def uncompressed_istft(spec: torch.Tensor, n_fft: int = N_FFT, hop: int = HOP,
                       window: Optional[torch.Tensor] = None) -> torch.Tensor:
    """complex64 (B, F, T) -> (B, hop*(T-1)).  core/function.py:695-703:
    decompress, irfft, synthesis window, overlap-add, divide by the squared
    window envelope, trim n_fft/2 at both ends (torch.istft, center=True)."""
    w = hamming_periodic(n_fft).to(spec.device) if window is None else window
    z = power_uncompress(spec)
    B, _, T = z.shape
    fr = torch.fft.irfft(z.transpose(1, 2), n=n_fft, dim=-1) * w      # (B, T, n_fft)
    full = n_fft + hop * (T - 1)
    y = torch.zeros(B, full, dtype=fr.dtype, device=fr.device)
    env = torch.zeros(full, dtype=fr.dtype, device=fr.device)
    w2 = w * w
    if fr.is_cuda:       # what torch.istft does: one col2im (bench.py's informational GPU-eager leg; a T-iteration loop would be launch-bound)
        y = F.fold(fr.transpose(1, 2), (1, full), (1, n_fft), stride=(1, hop)).reshape(B, full)
        env = F.fold(w2.expand(1, T, n_fft).transpose(1, 2), (1, full), (1, n_fft), stride=(1, hop)).reshape(full)
    else:
        for t in range(T):
            y[:, t * hop:t * hop + n_fft] += fr[:, t]
            env[t * hop:t * hop + n_fft] += w2
    half = n_fft // 2
    return y[:, half:full - half] / env[half:full - half]


# ----------------------------------------------------------------------------
# Generator building blocks (reference layout: (B, C, T, F))
# ----------------------------------------------------------------------------
def _inorm(x, sd, p):
    """nn.InstanceNorm2d(C, affine=True), eps 1e-5, biased plane variance."""
    return F.instance_norm(x, weight=sd[p + ".weight"], bias=sd[p + ".bias"], eps=1e-5)


def _prelu(x, sd, p):
    return F.prelu(x, sd[p + ".weight"])


def dilated_dense(x, sd: SD, p: str, depth: int = 4):
    """DilatedDenseNet.forward  models/generator.py:24-32."""
    skip = x
    out = x
    for i in range(1, depth + 1):
        dil = 2 ** (i - 1)
        h = F.pad(skip, (1, 1, dil, 0))
        h = F.conv2d(h, sd[f"{p}.conv{i}.weight"], sd[f"{p}.conv{i}.bias"], dilation=(dil, 1))
        h = _inorm(h, sd, f"{p}.norm{i}")
        out = _prelu(h, sd, f"{p}.prelu{i}")
        skip = torch.cat([out, skip], dim=1)
    return out


def dense_encoder(x_in, sd: SD, p: str = "dense_encoder"):
    """DenseEncoder.forward  models/generator.py:50-54."""
    h = F.conv2d(x_in, sd[p + ".conv_1.0.weight"], sd[p + ".conv_1.0.bias"])
    h = _prelu(_inorm(h, sd, p + ".conv_1.1"), sd, p + ".conv_1.2")
    h = dilated_dense(h, sd, p + ".dilated_dense")
    h = F.conv2d(h, sd[p + ".conv_2.0.weight"], sd[p + ".conv_2.0.bias"], stride=(1, 2), padding=(0, 1))
    return _prelu(_inorm(h, sd, p + ".conv_2.1"), sd, p + ".conv_2.2")


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _swish(x):
    return x * torch.sigmoid(x)


class TrainCtx:
    """Train-mode semantics of the generator (core/function.py:218-229 runs model(noisy_spec) under model.train()):
    * ``masks[site]``: the keep-mask (bool / 0-1, the shape of the dropped tensor) of every nn.Dropout(p=0.2) on the path --
      ``<conformer>.ff{1,2}.drop1`` after the Swish (conformer.py:139), ``.drop2`` after the second Linear (:141), ``<conformer>.attn.drop``
      on the projected attention output (:125); kept values are scaled by 1 / (1 - p) as nn.Dropout does.  The masks are INJECTED
      so that the reference, this restatement and the CUDA path see the same draw.
    * BatchNorm1d (conformer.py:167) uses batch statistics (biased variance) and updates ``running[<conv>.net.5.running_mean / _var]``
      with momentum 0.1 and the UNBIASED variance, ``num_batches_tracked`` += 1 -- collected in ``running`` instead of mutating sd."""

    def __init__(self, masks, p: float = 0.2):
        self.masks, self.p, self.running = masks, p, {}

    def drop(self, x, site):
        return x * self.masks[site].to(x.dtype) * (1.0 / (1.0 - self.p))


def feed_forward(x, sd: SD, p: str, tr: Optional[TrainCtx] = None):
    """Scale(0.5, PreNorm(FeedForward))  models/conformer.py:53-71,128-145 (eval: dropout off; train: two dropouts)."""
    h = _ln(x, sd, p + ".fn.norm")
    h = F.linear(h, sd[p + ".fn.fn.net.0.weight"], sd[p + ".fn.fn.net.0.bias"])
    h = _swish(h)
    if tr is not None:
        h = tr.drop(h, p + ".drop1")
    h = F.linear(h, sd[p + ".fn.fn.net.3.weight"], sd[p + ".fn.fn.net.3.bias"])
    if tr is not None:
        h = tr.drop(h, p + ".drop2")
    return 0.5 * h


def attention(x, sd: SD, p: str, chunk: int = 0, tr: Optional[TrainCtx] = None):
    """PreNorm(Attention) with Shaw relative positions  models/conformer.py:96-125.
    x: (S, n, 64).  ``chunk`` > 0 evaluates the S sequences in groups (the
    sequences are independent) to bound the (S, h, n, n) memory."""
    if tr is not None:
        return tr.drop(attention(x, sd, p, chunk), p + ".drop")       # conformer.py:125: dropout on the projected output only
    if chunk and x.shape[0] > chunk:
        return torch.cat([attention(x[i:i + chunk], sd, p) for i in range(0, x.shape[0], chunk)], 0)
    S, n, _ = x.shape
    h = _ln(x, sd, p + ".norm")
    q = F.linear(h, sd[p + ".fn.to_q.weight"])
    kv = F.linear(h, sd[p + ".fn.to_kv.weight"])
    k, v = kv[..., :HEADS * DIM_HEAD], kv[..., HEADS * DIM_HEAD:]
    split = lambda t: t.reshape(S, n, HEADS, DIM_HEAD).permute(0, 2, 1, 3)   # b h n d
    q, k, v = split(q), split(k), split(v)
    scale = DIM_HEAD ** -0.5
    dots = torch.matmul(q, k.transpose(-1, -2)) * scale
    pos = torch.arange(n, device=x.device)
    dist = (pos[:, None] - pos[None, :]).clamp(-MAX_POS, MAX_POS) + MAX_POS
    rel = sd[p + ".fn.rel_pos_emb.weight"][dist]                          # (n, n, d)
    dots = dots + torch.einsum("bhnd,nrd->bhnr", q, rel) * scale
    attn = dots.softmax(dim=-1)
    o = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(S, n, HEADS * DIM_HEAD)
    return F.linear(o, sd[p + ".fn.to_out.weight"], sd[p + ".fn.to_out.bias"])


def conv_module(x, sd: SD, p: str, tr: Optional[TrainCtx] = None):
    """ConformerConvModule  models/conformer.py:161-172 (eval: running statistics; train: batch statistics + running update)."""
    h = _ln(x, sd, p + ".net.0").transpose(1, 2)                           # b c n
    h = F.conv1d(h, sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])
    a, g = h.chunk(2, dim=1)
    h = a * torch.sigmoid(g)
    ks = sd[p + ".net.4.conv.weight"].shape[-1]
    h = F.pad(h, (ks // 2, ks // 2 - (ks + 1) % 2))
    h = F.conv1d(h, sd[p + ".net.4.conv.weight"], sd[p + ".net.4.conv.bias"], groups=h.shape[1])
    if tr is None:
        h = F.batch_norm(h, sd[p + ".net.5.running_mean"], sd[p + ".net.5.running_var"],
                         sd[p + ".net.5.weight"], sd[p + ".net.5.bias"], False, 0.0, 1e-5)
    else:
        rm, rv = sd[p + ".net.5.running_mean"].detach().clone(), sd[p + ".net.5.running_var"].detach().clone()
        h = F.batch_norm(h, rm, rv, sd[p + ".net.5.weight"], sd[p + ".net.5.bias"], True, 0.1, 1e-5)
        tr.running[p + ".net.5.running_mean"], tr.running[p + ".net.5.running_var"] = rm, rv
        tr.running[p + ".net.5.num_batches_tracked"] = sd[p + ".net.5.num_batches_tracked"] + 1
    h = _swish(h)
    h = F.conv1d(h, sd[p + ".net.7.weight"], sd[p + ".net.7.bias"])
    return h.transpose(1, 2)


def conformer_block(x, sd: SD, p: str, chunk: int = 0, tr: Optional[TrainCtx] = None):
    """ConformerBlock.forward  models/conformer.py:206-212."""
    x = feed_forward(x, sd, p + ".ff1", tr) + x
    x = attention(x, sd, p + ".attn", chunk, tr) + x
    x = conv_module(x, sd, p + ".conv", tr) + x
    x = feed_forward(x, sd, p + ".ff2", tr) + x
    return _ln(x, sd, p + ".post_norm")


def tscb(x, sd: SD, p: str, chunk: int = 0, tr: Optional[TrainCtx] = None):
    """TSCB.forward  models/generator.py:67-74.  x: (B, C, T, F)."""
    b, c, t, f = x.shape
    xt = x.permute(0, 3, 2, 1).reshape(b * f, t, c)
    xt = conformer_block(xt, sd, p + ".time_conformer", chunk, tr) + xt
    xf = xt.reshape(b, f, t, c).permute(0, 2, 1, 3).reshape(b * t, f, c)
    xf = conformer_block(xf, sd, p + ".freq_conformer", chunk, tr) + xf
    return xf.reshape(b, t, f, c).permute(0, 3, 1, 2)


def sub_pixel(x, sd: SD, p: str, r: int = 2):
    """SPConvTranspose2d.forward  models/generator.py:85-92."""
    y = F.conv2d(F.pad(x, (1, 1, 0, 0)), sd[p + ".conv.weight"], sd[p + ".conv.bias"])
    b, rc, t, w = y.shape
    y = y.reshape(b, r, rc // r, t, w).permute(0, 2, 3, 4, 1)
    return y.reshape(b, rc // r, t, w * r)


def mask_decoder(x, sd: SD, p: str = "mask_decoder"):
    """MaskDecoder.forward  models/generator.py:106-112 -> (B, 1, T, F)."""
    h = dilated_dense(x, sd, p + ".dense_block")
    h = sub_pixel(h, sd, p + ".sub_pixel")
    h = F.conv2d(h, sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"])
    h = _prelu(_inorm(h, sd, p + ".norm"), sd, p + ".prelu")
    h = F.conv2d(h, sd[p + ".final_conv.weight"], sd[p + ".final_conv.bias"])
    h = h.permute(0, 3, 2, 1).squeeze(-1)                                  # (B, F, T)
    h = F.prelu(h, sd[p + ".prelu_out.weight"])                            # slope per frequency bin
    return h.permute(0, 2, 1).unsqueeze(1)


def complex_decoder(x, sd: SD, p: str = "complex_decoder"):
    """ComplexDecoder.forward  models/generator.py:124-129 -> (B, 2, T, F)."""
    h = dilated_dense(x, sd, p + ".dense_block")
    h = sub_pixel(h, sd, p + ".sub_pixel")
    h = _prelu(_inorm(h, sd, p + ".norm"), sd, p + ".prelu")
    return F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"])


def tscnet_forward(spec: torch.Tensor, sd: SD, chunk: int = 0, stages: Optional[dict] = None, tr: Optional[TrainCtx] = None):
    """TSCNet.forward  models/generator.py:145-167.
    spec: complex64 (B, 201, T) -> (final_real, final_imag), each (B, 1, T, 201).
    If ``stages`` is a dict, per-stage tensors are stored in it (reference layout).  ``tr``: train-mode semantics (TrainCtx);
    InstanceNorm2d keeps no running statistics, so only the conformers differ between train and eval."""
    mag = spec.abs().unsqueeze(1).permute(0, 1, 3, 2)
    ph = spec.angle().unsqueeze(1).permute(0, 1, 3, 2)
    x_in = torch.cat([mag, spec.real.unsqueeze(1).permute(0, 1, 3, 2),
                      spec.imag.unsqueeze(1).permute(0, 1, 3, 2)], dim=1)
    h = dense_encoder(x_in, sd)
    if stages is not None:
        stages["x_in"] = x_in
        stages["encoder"] = h
    for i in range(1, 5):
        h = tscb(h, sd, f"TSCB_{i}", chunk, tr)
        if stages is not None:
            stages[f"tscb{i}"] = h
    mask = mask_decoder(h, sd)
    cplx = complex_decoder(h, sd)
    out_mag = mask * mag
    fr = out_mag * torch.cos(ph) + cplx[:, 0:1]
    fi = out_mag * torch.sin(ph) + cplx[:, 1:2]
    if stages is not None:
        stages["mask"] = mask
        stages["complex"] = cplx
    return fr, fi


# ----------------------------------------------------------------------------
# Diffusion variant: models/tsc_diffusion.py, models/DiffuSE.py:39-69
# ----------------------------------------------------------------------------
def diffusion_step_table(max_steps: int) -> torch.Tensor:
    """DiffusionEmbedding._build_embedding  models/DiffuSE.py:64-69 -> [max_steps, 128]."""
    steps = torch.arange(max_steps).unsqueeze(1)
    dims = torch.arange(64).unsqueeze(0)
    table = steps * 10.0 ** (dims * 4.0 / 63.0)
    return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


def diffusion_embedding(step: torch.Tensor, sd: SD, p: str, max_steps: int) -> torch.Tensor:
    """DiffusionEmbedding.forward  models/DiffuSE.py:46-62: table row (integer step) or the linear interpolation of the
    two neighbouring rows (fractional step), then Linear-SiLU-Linear-SiLU."""
    table = diffusion_step_table(max_steps)
    if step.dtype in (torch.int32, torch.int64):
        e = table[step]
    else:
        lo, hi = torch.floor(step).long(), torch.ceil(step).long()
        e = table[lo] + (table[hi] - table[lo]) * (step - lo).unsqueeze(-1)
    h = _swish(F.linear(e, sd[p + ".projection1.weight"], sd[p + ".projection1.bias"]))
    return _swish(F.linear(h, sd[p + ".projection2.weight"], sd[p + ".projection2.bias"]))


def merge_block(x, cond, step, sd: SD, max_steps: int, p: str = "merge_block"):
    """MergeBlock.forward  models/tsc_diffusion.py:27-41.  x, cond: (B, 64, T, F'); step: [1] or [B]."""
    d = diffusion_embedding(step, sd, p + ".diffusion_embedding", max_steps)
    d = F.linear(d, sd[p + ".diffusion_projection.weight"], sd[p + ".diffusion_projection.bias"])[:, :, None, None]
    c = F.conv2d(cond, sd[p + ".conditioner_projection.weight"], sd[p + ".conditioner_projection.bias"])
    y = F.conv2d(x + d, sd[p + ".merge_diffusion.weight"], sd[p + ".merge_diffusion.bias"]) + c
    gate, filt = torch.chunk(y, 2, dim=1)
    y = torch.sigmoid(gate) * torch.tanh(filt)
    r = F.conv2d(y, sd[p + ".output_residual.weight"], sd[p + ".output_residual.bias"])
    return (x + r) / math.sqrt(2.0)


def tsc_diffusion_forward(spec: torch.Tensor, noisy_spec: torch.Tensor, step: torch.Tensor, sd: SD, max_steps: int,
                          chunk: int = 0, stages: Optional[dict] = None):
    """tsc_diffusion.TSCNet.forward  models/tsc_diffusion.py:60-90.  spec, noisy_spec: complex64 (B, 201, T); the mask and
    the phase come from ``spec`` (the current estimate), ``noisy_spec`` only conditions the merge blocks."""
    def in3(z):
        return torch.cat([z.abs().unsqueeze(1).permute(0, 1, 3, 2), z.real.unsqueeze(1).permute(0, 1, 3, 2),
                          z.imag.unsqueeze(1).permute(0, 1, 3, 2)], dim=1)
    mag = spec.abs().unsqueeze(1).permute(0, 1, 3, 2)
    ph = spec.angle().unsqueeze(1).permute(0, 1, 3, 2)
    h = dense_encoder(in3(spec), sd, "dense_encoder")
    cond = dense_encoder(in3(noisy_spec), sd, "dense_encoder_noisy")
    if stages is not None:
        stages["encoder"], stages["encoder_noisy"] = h, cond
    for i in range(1, 5):
        h = tscb(merge_block(h, cond, step, sd, max_steps), sd, f"TSCB_{i}", chunk)
        if stages is not None:
            stages[f"tscb{i}"] = h
    mask = mask_decoder(h, sd)
    cplx = complex_decoder(h, sd)
    out_mag = mask * mag
    fr = out_mag * torch.cos(ph) + cplx[:, 0:1]
    fi = out_mag * torch.sin(ph) + cplx[:, 1:2]
    if stages is not None:
        stages["mask"], stages["complex"] = mask, cplx
    return fr, fi


def predict_tsc(wave: torch.Tensor, sd: SD, max_steps: int, T, c1, c2, c3, delta_bar, noises, chunk: int = 0,
                trace: Optional[list] = None) -> torch.Tensor:
    """predict_tsc  inference_diffuse.py:231-267, batched: (B, L) noisy -> (B, L) enhanced.  ``noises[n]`` is the (B, Lp)
    Gaussian draw the reference takes with torch.randn_like at step n > 0 (passed in so the run is reproducible)."""
    x = wave.to(torch.float32)
    B, L = x.shape
    c = torch.sqrt(L / torch.sum(x ** 2.0, dim=-1, keepdim=True))
    x = x * c
    pad = int(math.ceil(L / 100)) * 100 - L
    noisy_audio = torch.cat([x, x[:, :pad]], dim=-1)
    audio = noisy_audio
    orig = compressed_stft(noisy_audio)
    for n in range(len(c1) - 1, -1, -1):
        spec = compressed_stft(audio)
        fr, fi = tsc_diffusion_forward(spec, orig, torch.tensor([T[n]]), sd, max_steps, chunk)
        pred = uncompressed_istft(torch.complex(fr.permute(0, 1, 3, 2), fi.permute(0, 1, 3, 2)).squeeze(1))
        if n > 0:
            audio = c1[n] * audio + c2[n] * noisy_audio - c3[n] * pred
            audio = audio + delta_bar[n] ** 0.5 * noises[n]
        else:
            audio = c1[n] * audio - c3[n] * pred
            audio = (1 - 0.2) * audio + 0.2 * noisy_audio
        if trace is not None:
            trace.append(audio if n > 0 else audio / c)
    return (audio / c)[:, :L]


# ----------------------------------------------------------------------------
# wave -> wave  (inference_gan.py:75-100, batched)
# ----------------------------------------------------------------------------
def predict(wave: torch.Tensor, sd: SD, chunk: int = 0, stages: Optional[dict] = None) -> torch.Tensor:
    """(B, L) fp32 noisy -> (B, L) enhanced; every row gets predict()'s
    per-utterance semantics: RMS normalise, wrap-pad to a multiple of 100,
    compressed STFT, generator, decompress + iSTFT, un-normalise, trim."""
    x = wave.to(torch.float32)
    B, L = x.shape
    c = torch.sqrt(L / torch.sum(x ** 2.0, dim=-1, keepdim=True))
    x = x * c
    pad = int(math.ceil(L / 100)) * 100 - L
    x = torch.cat([x, x[:, :pad]], dim=-1)
    spec = compressed_stft(x)
    if stages is not None:
        stages["spec"] = spec
    fr, fi = tscnet_forward(spec, sd, chunk, stages)
    est = torch.complex(fr.permute(0, 1, 3, 2), fi.permute(0, 1, 3, 2)).squeeze(1)
    y = uncompressed_istft(est) / c
    return y[:, :L]


def si_sdr(est: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """Scale-invariant SDR in dB along the last axis (parity metric)."""
    est = est.double() - est.double().mean(-1, keepdim=True)
    ref = ref.double() - ref.double().mean(-1, keepdim=True)
    a = (est * ref).sum(-1, keepdim=True) / (ref * ref).sum(-1, keepdim=True)
    s = a * ref
    return 10.0 * torch.log10((s * s).sum(-1) / ((est - s) ** 2).sum(-1))
