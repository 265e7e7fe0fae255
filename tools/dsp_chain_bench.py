"""Consistency-loss DSP chain (core/function.py:227-254) forward + backward on libseb200 vs the reference's own torch.stft / torch.istft
autograd on the same GPU (cuFFT): ms per iteration at the training shape (4 x 2 s) and at 64 x 4 s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
import synth as weights  # noqa: E402

mse = torch.nn.functional.mse_loss


def ref_stft(x, win):      # core/function.py:685-693
    s = torch.stft(x, 400, 100, window=win, onesided=True, return_complex=True)
    mag, ph = s.abs() ** 0.3, s.angle()
    return torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))


def ref_istft(s, win):     # core/function.py:695-703
    mag, ph = s.abs() ** (1.0 / 0.3), s.angle()
    return torch.istft(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), 400, 100, window=win, onesided=True)


def chain(est, cspec, istft, stft):
    a = istft(est); p = stft(a); ca = istft(cspec); cp = stft(ca)
    return 0.9 * mse(p.abs(), cp.abs()) + 0.1 * (mse(p.real, cp.real) + mse(p.imag, cp.imag)) + 0.2 * torch.mean(torch.abs(a - ca))


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
win = torch.hamming_window(400, device="cuda")
for B, L in ((4, 32000), (64, 64000)):
    noisy, clean = weights.synth_wave(B, L, 7, "speech")
    with torch.no_grad():
        cspec = se_b200.compressed_stft(clean.cuda() * 3)
        est0 = se_b200.compressed_stft(noisy.cuda() * 3)

    def run(istft, stft):
        est = est0.clone().requires_grad_(True)
        chain(est, cspec, istft, stft).backward()
        return est.grad

    g1 = run(se_b200.uncompressed_istft, se_b200.compressed_stft)
    g2 = run(lambda s: ref_istft(s, win), lambda a: ref_stft(a, win))
    r = lambda t: torch.view_as_real(t)
    out[f"{B}x{L // 16000}s"] = {"seb200_ms": round(timed(lambda: run(se_b200.uncompressed_istft, se_b200.compressed_stft)), 3),
                                "torch_cufft_ms": round(timed(lambda: run(lambda s: ref_istft(s, win), lambda a: ref_stft(a, win))), 3),
                                "grad_rel_l2_vs_torch": float((r(g1) - r(g2)).norm() / r(g2).norm())}
print(json.dumps(out))
