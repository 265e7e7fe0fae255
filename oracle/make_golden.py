"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build
container, where /root/reference is mounted):

    python oracle/make_golden.py

Each file holds seeded inputs and what the reference's own code returns for
them: ``inference_gan.predict`` (wave -> wave), ``core.function.compressed_stft``
and ``models.generator.TSCNet.forward``, with weights from
``oracle.weights.synth_state_dict`` loaded through ``load_state_dict(strict)``.
The GPU box has no /root/reference; its tests read these files instead.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, weights  # noqa: E402

CASES = [
    # name, batch, length, wave seed, weight seed, kind
    ("speech_b2_L8000", 2, 8000, 1234, 0, "speech"),
    ("noise_b1_L4050_wrap", 1, 4050, 77, 1, "noise"),      # exercises predict()'s wrap-pad branch
    ("speech_b1_L16000", 1, 16000, 5, 0, "speech"),
]


def main():
    ref = ref_import.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, b, L, wseed, sseed, kind in CASES:
        sd = weights.synth_state_dict(sseed)
        model = ref.TSCNet(num_channel=64, num_features=201)
        model.load_state_dict(sd, strict=True)
        model.eval()
        noisy, clean = weights.synth_wave(b, L, wseed, kind)
        enhanced = np.stack([ref.predict(model, ref.config, noisy[i].numpy(), device=torch.device("cpu"))
                             for i in range(b)])
        # the pieces of predict(), for stage-level parity (inference_gan.py:79-90)
        with torch.no_grad():
            c = torch.sqrt(L / torch.sum(noisy ** 2.0, dim=-1, keepdim=True))
            x = noisy * c
            pad = int(np.ceil(L / 100)) * 100 - L
            x = torch.cat([x, x[:, :pad]], dim=-1)
            spec = ref.compressed_stft(x, 400, 100, torch.hamming_window(400))
            fr, fi = model(spec)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            noisy=noisy.numpy(), clean=clean.numpy(), enhanced=enhanced.astype(np.float32),
            spec_real=spec.real.numpy(), spec_imag=spec.imag.numpy(),
            final_real=fr.numpy(), final_imag=fi.numpy(),
            weight_seed=np.int64(sseed), wave_seed=np.int64(wseed))
        print(name, "enhanced peak", float(np.abs(enhanced).max()))
    diffusion_golden(ref, out_dir)


# models/tsc_diffusion.py:TSCNet.forward on (estimate, conditioning utterance, step); integer, fractional
# (DiffuSE.py:57-62 interpolation) and per-utterance steps
DIFFUSION_CASE = dict(name="diffusion_b2_L3000", batch=2, length=3000, wave_seed=11, weight_seed=3, max_steps=50,
                      steps=[("int1", [7], "int64"), ("frac1", [3.4], "float32"), ("intB", [2, 40], "int64")])


def diffusion_inputs(case=DIFFUSION_CASE):
    """(estimate wave, conditioning wave): the conditioning utterance is the RMS-normalised noisy clip; the estimate is a
    partly denoised mixture, as the reverse process holds mid-way (inference_diffuse.py:246-262)."""
    noisy, clean = weights.synth_wave(case["batch"], case["length"], case["wave_seed"], "speech")
    c = torch.sqrt(noisy.shape[-1] / torch.sum(noisy ** 2.0, dim=-1, keepdim=True))
    g = torch.Generator().manual_seed(case["wave_seed"] + 1)
    est = c * (0.5 * clean + 0.5 * noisy) + 0.05 * torch.randn(noisy.shape, generator=g)
    return est, c * noisy


def diffusion_golden(ref, out_dir):
    case = DIFFUSION_CASE
    sd = weights.synth_state_dict(case["weight_seed"], spec=weights.tsc_diffusion_spec())
    model = ref.DiffusionTSCNet(num_channel=64, num_features=201, noise_schedule=list(range(case["max_steps"])))
    model.load_state_dict(sd, strict=True)
    model.eval()
    est, cond = diffusion_inputs(case)
    out = {}
    with torch.no_grad():
        win = torch.hamming_window(400)
        sx, sn = ref.compressed_stft(est, 400, 100, win), ref.compressed_stft(cond, 400, 100, win)
        for tag, vals, dt in case["steps"]:
            fr, fi = model(sx, sn, torch.tensor(vals, dtype=getattr(torch, dt)))
            out[f"final_real_{tag}"], out[f"final_imag_{tag}"] = fr.numpy(), fi.numpy()
    np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"), est=est.numpy(), cond=cond.numpy(),
                        spec_x_real=sx.real.numpy(), spec_x_imag=sx.imag.numpy(), spec_n_real=sn.real.numpy(), spec_n_imag=sn.imag.numpy(),
                        weight_seed=np.int64(case["weight_seed"]), wave_seed=np.int64(case["wave_seed"]), max_steps=np.int64(case["max_steps"]), **out)
    print(case["name"], {k: float(np.abs(v).max()) for k, v in out.items()})


if __name__ == "__main__":
    main()
