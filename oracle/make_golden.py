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
    default_init_golden(ref, out_dir)
    diffusion_golden(ref, out_dir)


def default_init_golden(ref, out_dir):
    """second weight set: PyTorch-default initialisation (oracle.weights.torch_default_state_dict)"""
    name, b, L, wseed, sseed = "default_init_b1_L6000", 1, 6000, 55, 2
    sd = weights.torch_default_state_dict(sseed)
    model = ref.TSCNet(num_channel=64, num_features=201)
    model.load_state_dict(sd, strict=True)
    model.eval()
    noisy, clean = weights.synth_wave(b, L, wseed, "speech")
    enhanced = np.stack([ref.predict(model, ref.config, noisy[i].numpy(), device=torch.device("cpu")) for i in range(b)])
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), noisy=noisy.numpy(), clean=clean.numpy(), enhanced=enhanced.astype(np.float32),
                        weight_seed=np.int64(sseed), wave_seed=np.int64(wseed))
    print(name, "enhanced peak", float(np.abs(enhanced).max()))


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
    np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"),
                        spec_x_real=sx.real.numpy(), spec_x_imag=sx.imag.numpy(), spec_n_real=sn.real.numpy(), spec_n_imag=sn.imag.numpy(),
                        weight_seed=np.int64(case["weight_seed"]), wave_seed=np.int64(case["wave_seed"]), max_steps=np.int64(case["max_steps"]), **out)
    print(case["name"], {k: float(np.abs(v).max()) for k, v in out.items()})
    reverse_golden(ref, out_dir, sd, model)


REVERSE_CASE = dict(name="diffusion_reverse_L2950", length=2950, wave_seed=31, noise_seed=123)     # 2950: predict_tsc's wrap-pad branch


def reverse_golden(ref, out_dir, sd, model):
    """the reference's own predict_tsc (inference_diffuse.py:231-267) with its fast-sampling schedule (inference_schedule on
    config/default.py:27-28,119: 50 training steps, 6 inference steps); torch.manual_seed pins its randn_like draws."""
    import types
    case = REVERSE_CASE
    cfg = types.SimpleNamespace(N_FFT=400, HOP_SAMPLES=100, NOISE_SCHEDULE=np.linspace(1e-4, 0.035, DIFFUSION_CASE["max_steps"]).tolist(),
                                INFERENCE_NOISE_SCHEDULE=[0.0001, 0.001, 0.01, 0.05, 0.2, 0.35])
    alpha, beta, alpha_cum, sigmas, T, c1, c2, c3, delta, delta_bar = ref.inference_schedule(cfg, fast_sampling=True)
    noisy, _ = weights.synth_wave(1, case["length"], case["wave_seed"], "speech")
    torch.manual_seed(case["noise_seed"])
    y = ref.predict_tsc(model, types.SimpleNamespace(comp_type="pow"), cfg, noisy[0].numpy(), alpha, beta, alpha_cum, sigmas, T, c1, c2, c3,
                        delta, delta_bar, device=torch.device("cpu"))
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"), noisy=noisy.numpy(), enhanced=y.astype(np.float32)[None],
                        T=np.asarray(T, dtype=np.float32), c1=f64(c1), c2=f64(c2), c3=f64(c3), delta_bar=f64(delta_bar), alpha=f64(alpha),
                        weight_seed=np.int64(DIFFUSION_CASE["weight_seed"]), max_steps=np.int64(DIFFUSION_CASE["max_steps"]),
                        noise_seed=np.int64(case["noise_seed"]), wave_seed=np.int64(case["wave_seed"]))
    print(case["name"], "T", T, "enhanced peak", float(np.abs(y).max()))


if __name__ == "__main__":
    main()
