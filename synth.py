"""Deterministic synthetic weights and waveforms for the generator (inputs of bench.py, the tests and the golden scripts;
not part of the oracle: nothing here restates the reference's arithmetic).

``tscnet_spec()`` lists the 359 ``state_dict`` entries of the reference's
``TSCNet(num_channel=64, num_features=201)`` (models/generator.py:132-143) in
registration order; ``synth_state_dict(seed)`` fills them with values that do
not depend on torch's module-construction RNG order, so the same tensors can
be rebuilt on any box: every entry draws from its own generator seeded by
(seed, index).  Conv/Linear weights are Kaiming-normal (std = sqrt(2/fan_in),
as utils/utils.py:92-104 applies them), biases 0.01 + noise, embeddings N(0,1);
norm affine parameters, PReLU slopes and BatchNorm running statistics are
perturbed away from their defaults so that every affine path is exercised.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def _dense(p, kinds):
    for i in range(1, 5):
        kinds += [(f"{p}.conv{i}.weight", (64, 64 * i, 2, 3), "w"), (f"{p}.conv{i}.bias", (64,), "b"),
                  (f"{p}.norm{i}.weight", (64,), "g"), (f"{p}.norm{i}.bias", (64,), "beta"),
                  (f"{p}.prelu{i}.weight", (64,), "slope")]


def _ff(p, kinds):
    kinds += [(f"{p}.fn.fn.net.0.weight", (256, 64), "w"), (f"{p}.fn.fn.net.0.bias", (256,), "b"),
              (f"{p}.fn.fn.net.3.weight", (64, 256), "w"), (f"{p}.fn.fn.net.3.bias", (64,), "b"),
              (f"{p}.fn.norm.weight", (64,), "g"), (f"{p}.fn.norm.bias", (64,), "beta")]


def _conformer(p, kinds):
    _ff(p + ".ff1", kinds)
    kinds += [(f"{p}.attn.fn.to_q.weight", (64, 64), "w"), (f"{p}.attn.fn.to_kv.weight", (128, 64), "w"),
              (f"{p}.attn.fn.to_out.weight", (64, 64), "w"), (f"{p}.attn.fn.to_out.bias", (64,), "b"),
              (f"{p}.attn.fn.rel_pos_emb.weight", (1025, 16), "emb"),
              (f"{p}.attn.norm.weight", (64,), "g"), (f"{p}.attn.norm.bias", (64,), "beta"),
              (f"{p}.conv.net.0.weight", (64,), "g"), (f"{p}.conv.net.0.bias", (64,), "beta"),
              (f"{p}.conv.net.2.weight", (256, 64, 1), "w"), (f"{p}.conv.net.2.bias", (256,), "b"),
              (f"{p}.conv.net.4.conv.weight", (128, 1, 31), "w"), (f"{p}.conv.net.4.conv.bias", (128,), "b"),
              (f"{p}.conv.net.5.weight", (128,), "g"), (f"{p}.conv.net.5.bias", (128,), "beta"),
              (f"{p}.conv.net.5.running_mean", (128,), "rm"), (f"{p}.conv.net.5.running_var", (128,), "rv"),
              (f"{p}.conv.net.5.num_batches_tracked", (), "count"),
              (f"{p}.conv.net.7.weight", (64, 128, 1), "w"), (f"{p}.conv.net.7.bias", (64,), "b")]
    _ff(p + ".ff2", kinds)
    kinds += [(f"{p}.post_norm.weight", (64,), "g"), (f"{p}.post_norm.bias", (64,), "beta")]


def tscnet_spec():
    """[(key, shape, kind)] in the reference's state_dict order."""
    k = []
    e = "dense_encoder"
    k += [(f"{e}.conv_1.0.weight", (64, 3, 1, 1), "w"), (f"{e}.conv_1.0.bias", (64,), "b"),
          (f"{e}.conv_1.1.weight", (64,), "g"), (f"{e}.conv_1.1.bias", (64,), "beta"),
          (f"{e}.conv_1.2.weight", (64,), "slope")]
    _dense(f"{e}.dilated_dense", k)
    k += [(f"{e}.conv_2.0.weight", (64, 64, 1, 3), "w"), (f"{e}.conv_2.0.bias", (64,), "b"),
          (f"{e}.conv_2.1.weight", (64,), "g"), (f"{e}.conv_2.1.bias", (64,), "beta"),
          (f"{e}.conv_2.2.weight", (64,), "slope")]
    for i in range(1, 5):
        _conformer(f"TSCB_{i}.time_conformer", k)
        _conformer(f"TSCB_{i}.freq_conformer", k)
    m = "mask_decoder"
    _dense(f"{m}.dense_block", k)
    k += [(f"{m}.sub_pixel.conv.weight", (128, 64, 1, 3), "w"), (f"{m}.sub_pixel.conv.bias", (128,), "b"),
          (f"{m}.conv_1.weight", (1, 64, 1, 2), "w"), (f"{m}.conv_1.bias", (1,), "b"),
          (f"{m}.norm.weight", (1,), "g"), (f"{m}.norm.bias", (1,), "beta"),
          (f"{m}.prelu.weight", (1,), "slope"),
          (f"{m}.final_conv.weight", (1, 1, 1, 1), "w1"), (f"{m}.final_conv.bias", (1,), "b"),
          (f"{m}.prelu_out.weight", (201,), "slope_neg")]
    c = "complex_decoder"
    _dense(f"{c}.dense_block", k)
    k += [(f"{c}.sub_pixel.conv.weight", (128, 64, 1, 3), "w"), (f"{c}.sub_pixel.conv.bias", (128,), "b"),
          (f"{c}.prelu.weight", (64,), "slope"),
          (f"{c}.norm.weight", (64,), "g"), (f"{c}.norm.bias", (64,), "beta"),
          (f"{c}.conv.weight", (2, 64, 1, 2), "w"), (f"{c}.conv.bias", (2,), "b")]
    return k


def tsc_diffusion_spec():
    """[(key, shape, kind)] of models/tsc_diffusion.py:TSCNet in the reference's state_dict order: the generator's entries with
    ``dense_encoder_noisy`` and ``merge_block`` inserted after ``dense_encoder`` (tsc_diffusion.py:46-57)."""
    base = tscnet_spec()
    enc = [e for e in base if e[0].startswith("dense_encoder.")]
    rest = [e for e in base if not e[0].startswith("dense_encoder.")]
    noisy = [(k.replace("dense_encoder.", "dense_encoder_noisy.", 1), sh, kind) for k, sh, kind in enc]
    m = "merge_block"
    merge = [(f"{m}.diffusion_embedding.projection1.weight", (512, 128), "w"), (f"{m}.diffusion_embedding.projection1.bias", (512,), "b"),
             (f"{m}.diffusion_embedding.projection2.weight", (512, 512), "w"), (f"{m}.diffusion_embedding.projection2.bias", (512,), "b"),
             (f"{m}.diffusion_projection.weight", (64, 512), "w"), (f"{m}.diffusion_projection.bias", (64,), "b"),
             (f"{m}.merge_diffusion.weight", (128, 64, 1, 1), "w"), (f"{m}.merge_diffusion.bias", (128,), "b"),
             (f"{m}.conditioner_projection.weight", (128, 64, 1, 1), "w"), (f"{m}.conditioner_projection.bias", (128,), "b"),
             (f"{m}.output_residual.weight", (64, 64, 1, 1), "w"), (f"{m}.output_residual.bias", (64,), "b")]
    return enc + noisy + merge + rest


def torch_default_state_dict(seed: int = 0, spec=None) -> "OrderedDict[str, torch.Tensor]":
    """The second weight set of SURVEY 8d: PyTorch's DEFAULT initialisation of every layer (what ``TSCNet(...)`` holds before
    ``model.apply(kaiming_init)``): Linear / ConvNd weights kaiming_uniform(a = sqrt 5), i.e. U(-1/sqrt(fan_in), 1/sqrt(fan_in)), biases
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) with the fan-in of their layer, Embedding N(0, 1), norms 1 / 0, BatchNorm running statistics 0 / 1,
    PReLU 0.25 (``prelu_out`` -0.25, generator.py:104).  Drawn from per-entry generators so any box rebuilds the same tensors."""
    sd = OrderedDict()
    fan_in = 1
    for idx, (key, shape, kind) in enumerate(tscnet_spec() if spec is None else spec):
        g = torch.Generator().manual_seed(7919 + seed * 100003 + idx)
        uni = lambda bound: (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        if kind in ("w", "w1"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = uni(1.0 / math.sqrt(fan_in))
        elif kind == "b":
            t = uni(1.0 / math.sqrt(fan_in))          # the bias entry follows its layer's weight in the spec
        elif kind in ("g", "rv"):
            t = torch.ones(shape)
        elif kind in ("beta", "rm"):
            t = torch.zeros(shape)
        elif kind == "slope":
            t = torch.full(shape, 0.25)
        elif kind == "slope_neg":
            t = torch.full(shape, -0.25)
        elif kind == "emb":
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.int64)
        else:
            raise KeyError(kind)
        sd[key] = t
    return sd


def synth_state_dict(seed: int = 0, perturb: float = 0.1, spec=None) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for idx, (key, shape, kind) in enumerate(tscnet_spec() if spec is None else spec):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        rn = lambda: torch.randn(shape, generator=g, dtype=torch.float32)
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = rn() * math.sqrt(2.0 / fan_in)
        elif kind == "w1":
            t = 1.0 + perturb * rn()
        elif kind == "b":
            t = 0.01 + 0.5 * perturb * rn()
        elif kind == "g":
            t = 1.0 + perturb * rn()
        elif kind == "beta":
            t = perturb * rn()
        elif kind == "slope":
            t = 0.25 + 0.5 * perturb * rn()
        elif kind == "slope_neg":
            t = -0.25 + 0.5 * perturb * rn()
        elif kind == "emb":
            t = rn()
        elif kind == "rm":
            t = perturb * rn()
        elif kind == "rv":
            t = 1.0 + perturb * torch.rand(shape, generator=g)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.int64)
        else:
            raise KeyError(kind)
        sd[key] = t
    return sd


def synth_wave(batch: int, length: int, seed: int = 1234, kind: str = "speech"):
    """Synthetic 16 kHz utterances.  'noise': white Gaussian sigma 0.1.
    'speech': amplitude-modulated harmonic + noise; returns (noisy, clean)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length, dtype=torch.float64) / 16000.0
    if kind == "noise":
        x = 0.1 * torch.randn(batch, length, generator=g, dtype=torch.float32)
        return x, x
    f0 = 100.0 + 200.0 * torch.rand(batch, 1, generator=g, dtype=torch.float64)
    clean = (0.3 * torch.sin(2 * math.pi * f0 * t) * (0.5 + 0.5 * torch.sin(2 * math.pi * 3.0 * t))
             + 0.1 * torch.sin(2 * math.pi * 8.0 * f0 * t)).to(torch.float32)
    clean = clean + 0.01 * torch.randn(batch, length, generator=g, dtype=torch.float32)
    noisy = clean + 0.1 * torch.randn(batch, length, generator=g, dtype=torch.float32)
    return noisy, clean


# ---- train-mode inputs: dropout keep-masks and output cotangents, reproducible on any box ------------------------
def dropout_sites():
    """[(site, channels)] of the 40 active nn.Dropout(p=0.2) modules of TSCNet in forward order (conformer.py:125,139,141; the conv
    module's Dropout(0.0) at :172 is the identity): per conformer ff1.drop1 (256, after Swish), ff1.drop2 (64), attn.drop (64),
    ff2.drop1 (256), ff2.drop2 (64)."""
    sites = []
    for i in range(1, 5):
        for ax in ("time", "freq"):
            p = f"TSCB_{i}.{ax}_conformer"
            sites += [(f"{p}.ff1.drop1", 256), (f"{p}.ff1.drop2", 64), (f"{p}.attn.drop", 64), (f"{p}.ff2.drop1", 256), (f"{p}.ff2.drop2", 64)]
    return sites


def dropout_masks(seed: int, B: int, T: int, Fh: int, p: float = 0.2):
    """site -> bool keep-mask [B, T, Fh, C] in the channels-last token order of the CUDA path (one CPU generator per site)."""
    out = {}
    for idx, (site, ch) in enumerate(dropout_sites()):
        g = torch.Generator().manual_seed(424243 + seed * 1009 + idx)
        out[site] = torch.rand(B, T, Fh, ch, generator=g) >= p
    return out


def masks_reference_layout(masks):
    """the same masks in the layout each reference module sees (generator.py:68-71): time conformer (B*F', T, C), frequency conformer (B*T, F', C)"""
    out = {}
    for site, m in masks.items():
        B, T, Fh, C = m.shape
        out[site] = m.permute(0, 2, 1, 3).reshape(B * Fh, T, C) if ".time_conformer." in site else m.reshape(B * T, Fh, C)
    return out


def cotangents(seed: int, B: int, T: int, F: int = 201):
    """fixed output cotangents (d loss / d final_real, d loss / d final_imag), each (B, 1, T, F): the scalar the gradient tests
    differentiate is sum(final_real * g_real + final_imag * g_imag)"""
    g = torch.Generator().manual_seed(90001 + seed)
    return torch.randn(B, 1, T, F, generator=g), torch.randn(B, 1, T, F, generator=g)


def has_zero_gradient(key: str) -> bool:
    """biases that sit directly in front of a normalisation which removes a per-plane / per-batch constant (Conv2d -> InstanceNorm2d:
    generator.py:18-22,39-40,45-46,100-101; depthwise Conv1d -> BatchNorm1d in train mode: conformer.py:166-167): their true gradient is
    exactly zero, autograd returns rounding noise, so gradient tests bound their magnitude instead of a relative error"""
    import re
    return bool(re.search(r"(dilated_dense|dense_block)\.conv[1-4]\.bias$|conv_1\.0\.bias$|conv_2\.0\.bias$|mask_decoder\.conv_1\.bias$|net\.4\.conv\.bias$", key))
