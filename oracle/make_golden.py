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


if __name__ == "__main__":
    main()
