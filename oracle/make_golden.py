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
    train_golden(ref, out_dir)


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


# ---- long clips (BASELINE configs[2]: 30 s, T = 4801; and 10 s, T = 1001, both clamp sides active inside one sequence) --------
LONG_CASES = [
    # name, length, wave seed, weight seed
    ("long_b1_L100000", 100000, 21, 0),
    ("long_b1_L480000", 480000, 99, 1),
]


def _chunked_tscb_forward(self, x_in, chunk=8):
    """TSCB.forward (generator.py:67-74) with the reference's OWN time / frequency ConformerBlock modules driven over chunks of
    <= `chunk` sequences: in eval mode every (b, f) / (b, t) sequence is independent through a ConformerBlock (BatchNorm uses
    running statistics; tests/test_oracle.py checks this), and the un-chunked call would materialise 3 x (101, 4, 4801, 4801)
    score tensors (110 GB) at 30 s."""
    b, c, t, f = x_in.size()
    x_t = x_in.permute(0, 3, 2, 1).contiguous().view(b * f, t, c)
    x_t = torch.cat([self.time_conformer(x_t[i:i + chunk]) for i in range(0, b * f, chunk)], dim=0) + x_t
    x_f = x_t.view(b, f, t, c).permute(0, 2, 1, 3).contiguous().view(b * t, f, c)
    x_f = torch.cat([self.freq_conformer(x_f[i:i + 256]) for i in range(0, b * t, 256)], dim=0) + x_f
    return x_f.view(b, t, f, c).permute(0, 3, 1, 2)


def long_golden(ref, out_dir, only=None):
    """the reference's predict() on one long clip; only the enhanced waveform is stored (the input is re-made from the seeds)"""
    import time
    import types
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, L, wseed, sseed in LONG_CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        model = ref.TSCNet(num_channel=64, num_features=201)
        model.load_state_dict(weights.synth_state_dict(sseed), strict=True)
        model.eval()
        for k in range(1, 5):
            blk = getattr(model, f"TSCB_{k}")
            blk.forward = types.MethodType(_chunked_tscb_forward, blk)
        noisy, _ = weights.synth_wave(1, L, wseed, "speech")
        enhanced = ref.predict(model, ref.config, noisy[0].numpy(), device=torch.device("cpu"))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), enhanced=enhanced.astype(np.float32)[None],
                            weight_seed=np.int64(sseed), wave_seed=np.int64(wseed), length=np.int64(L))
        print(name, "enhanced peak", float(np.abs(enhanced).max()), f"{time.time() - t0:.0f} s", flush=True)


# ---- train mode: the reference's TSCNet under .train() with its nn.Dropout modules patched to fixed masks ----------
TRAIN_CASE = dict(name="train_b2_L10000", batch=2, length=10000, wave_seed=41, weight_seed=4, mask_seed=7, cot_seed=1)


def _patch_dropouts(model, masks_ref, p=0.2):
    """replace the forward of every active nn.Dropout of the reference model by x * mask / (1 - p) with the injected mask"""
    name_of = {}
    for i in range(1, 5):
        for ax in ("time", "freq"):
            q = f"TSCB_{i}.{ax}_conformer"
            name_of[f"{q}.ff1.fn.fn.net.2"] = f"{q}.ff1.drop1"; name_of[f"{q}.ff1.fn.fn.net.4"] = f"{q}.ff1.drop2"
            name_of[f"{q}.ff2.fn.fn.net.2"] = f"{q}.ff2.drop1"; name_of[f"{q}.ff2.fn.fn.net.4"] = f"{q}.ff2.drop2"
            name_of[f"{q}.attn.fn.dropout"] = f"{q}.attn.drop"
    seen = 0
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Dropout) and mod.p > 0:
            site = name_of[name]
            mod.forward = (lambda m: (lambda x: x * m.to(x.dtype) * (1.0 / (1.0 - p))))(masks_ref[site])
            seen += 1
    assert seen == len(name_of) == 40, seen


def train_golden(ref, out_dir):
    """The unmodified reference TSCNet under .train() with injected dropout masks, twice: in float32 (what main_gan.py runs) and in
    float64 (`model.double()`, the same modules: the truth the float32 run is an approximation of).  Stored: the compressed spectrogram
    fed to the model (every implementation under test starts from these bits), the train-mode outputs and the BatchNorm buffers after
    the step (float32 run), the gradient of sum(final_real * g_r + final_imag * g_i) for all 335 parameters from the FLOAT64 run
    (rounded to float32), and per parameter the rel-L2 distance of the reference's own float32 gradient from it (`ref32_err:`).
    Why both: this random-weight network amplifies perturbations ~100x from input to gradient (a 1e-5 change of the spectrogram moves
    gradients by 1e-3), so two correct float32 implementations differ by a few 1e-3; the float64 run is the common yardstick."""
    c = TRAIN_CASE
    sd = weights.synth_state_dict(c["weight_seed"])
    noisy, _ = weights.synth_wave(c["batch"], c["length"], c["wave_seed"], "speech")
    cfac = torch.sqrt(noisy.shape[-1] / torch.sum(noisy ** 2.0, dim=-1, keepdim=True))
    spec = ref.compressed_stft(noisy * cfac, 400, 100, torch.hamming_window(400))
    B, _, T = spec.shape
    masks = weights.masks_reference_layout(weights.dropout_masks(c["mask_seed"], B, T, 101))
    gr, gi = weights.cotangents(c["cot_seed"], B, T)

    def run(dt):
        model = ref.TSCNet(num_channel=64, num_features=201)
        model.load_state_dict(sd, strict=True)
        model.train()
        model = model.to(dt)
        _patch_dropouts(model, masks)
        fr, fi = model(spec.to(torch.complex128 if dt == torch.float64 else torch.complex64))
        ((fr * gr.to(dt)).sum() + (fi * gi.to(dt)).sum()).backward()
        return model, fr.detach(), fi.detach()

    m32, fr, fi = run(torch.float32)
    m64, fr64, _ = run(torch.float64)
    out = {"spec_real": spec.real.numpy(), "spec_imag": spec.imag.numpy(), "final_real": fr.numpy(), "final_imag": fi.numpy()}
    for k, v in m32.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["buf:" + k] = v.numpy()
    p32 = dict(m32.named_parameters())
    worst = 0.0
    for k, prm in m64.named_parameters():
        t = prm.grad
        out["grad:" + k] = t.to(torch.float32).numpy()
        e = float((p32[k].grad.double() - t).norm() / t.norm().clamp_min(1e-300))
        out["ref32_err:" + k] = np.float32(e)
        if not weights.has_zero_gradient(k):
            worst = max(worst, e)
    np.savez_compressed(os.path.join(out_dir, c["name"] + ".npz"), **out, **{k: np.int64(v) for k, v in c.items() if k != "name"})
    print(c["name"], "335 gradients; reference float32 vs float64: worst rel-L2", worst, "forward", float((fr.double() - fr64).abs().max() / fr64.abs().max()))


if __name__ == "__main__":
    if "--train" in sys.argv:
        train_golden(ref_import.load(), os.path.join(ROOT, "tests", "golden"))
    elif "--long" in sys.argv:
        long_golden(ref_import.load(), os.path.join(ROOT, "tests", "golden"), only=[a for a in sys.argv[1:] if not a.startswith("--")])
    else:
        main()
