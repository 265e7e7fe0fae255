"""Reverse-diffusion enhancement with the diffusion TSCNet: drop-in for ``predict_tsc`` (inference_diffuse.py:231-267).

    y = predict_tsc(model, args, config, noisy_signal, alpha, beta, alpha_cum, sigmas, T, c1, c2, c3, delta, delta_bar)

keeps the reference's argument list (the schedule arrays are what its ``inference_schedule`` returns; that function is
host-side numpy and is not rebuilt here).  ``DiffusionEnhancerB200.reverse`` is the batched form.  Per step *n* = N-1 .. 0:

    spec_n = compressed_stft(audio)                              STFT engine, writes the network's 3-channel input directly
    est    = model(spec_n, spec(noisy), T[n])                    tsc_diffusion.TSCNet on libseb200
    pred   = uncompressed_istft(est)                             decompress + iDFT + overlap-add
    audio  = c1[n] audio + c2[n] noisy - c3[n] pred + sqrt(delta_bar[n]) N(0, 1)         (n > 0)
    audio  = (1 - gamma) (c1[0] audio - c3[0] pred) + gamma noisy,  gamma = 0.2            (n = 0)

and the per-utterance RMS gain is removed at the end.  The conditioning spectrogram is computed once; nothing but the
waveform update's Gaussian noise (``torch.randn`` on the device, or the caller's ``noise_fn``) comes from PyTorch.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch

from . import dsp, ops
from .tsc_diffusion import TSCNet

GAMMA0 = 0.2          # inference_diffuse.py:245 (gamma = [0.2], only gamma[0] is read)


class DiffusionEnhancerB200:
    def __init__(self, model: TSCNet):
        self.model = model
        self.dft_engine = dsp.DFT_ENGINE

    @torch.no_grad()
    def reverse(self, noisy: torch.Tensor, T, c1, c2, c3, delta_bar, noise_fn: Optional[Callable[[int, tuple], torch.Tensor]] = None,
                trace: Optional[list] = None) -> torch.Tensor:
        """noisy: (B, L) fp32 CUDA -> enhanced (B, L).  ``noise_fn(n, shape)`` supplies the step-n Gaussian noise (B, Lp) on the
        device (default: torch.randn); ``trace`` collects the waveform after every step (tests)."""
        if not noisy.is_cuda:
            raise RuntimeError("DiffusionEnhancerB200 has no CPU path: move the waveform to the GPU first")
        x = noisy.to(torch.float32).contiguous()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        B, L = x.shape
        nsteps = len(c1)
        if not (len(T) == len(c2) == len(c3) == len(delta_bar) == nsteps) or nsteps < 1:
            raise RuntimeError("schedule arrays T, c1, c2, c3, delta_bar must have one entry per inference step")
        with torch.cuda.device(x.device):                 # launches follow the input's device
            noisy_audio, cond_in3, c = self.prepare(x)
            audio = noisy_audio.contiguous()
            for n in range(nsteps - 1, -1, -1):
                noise = None
                if n > 0:
                    noise = noise_fn(n, tuple(audio.shape)) if noise_fn is not None else torch.randn(audio.shape, device=x.device)
                audio = self.step(audio, noisy_audio, cond_in3, n, T, c1, c2, c3, delta_bar, noise, c)
                if trace is not None:
                    trace.append(audio.clone())
        return audio[:, :L]


    @torch.no_grad()
    def prepare(self, x: torch.Tensor):
        """(B, L) fp32 CUDA -> (noisy_audio (B, Lp) view, conditioning in3 [B, T, 201, 3], gain c [B]): RMS-normalise, wrap-pad to a
        multiple of 100, compressed STFT of the conditioning utterance (inference_diffuse.py:235-247)."""
        B, L = x.shape
        Lp = int(math.ceil(L / dsp.HOP)) * dsp.HOP
        xpad, c = ops.rms_pad(x, Lp, normalize=True)
        noisy_audio = xpad[:, dsp.N_FFT // 2:dsp.N_FFT // 2 + Lp]          # row-strided view of the padded buffer
        return noisy_audio, dsp.stft_in3(xpad, Lp // dsp.HOP + 1, self.dft_engine), c

    @torch.no_grad()
    def step(self, audio, noisy_audio, cond_in3, n: int, T, c1, c2, c3, delta_bar, noise=None, c=None) -> torch.Tensor:
        """one reverse step n (inference_diffuse.py:248-264): STFT(audio) -> network at step T[n] -> iSTFT -> waveform update.
        n > 0 needs ``noise`` (B, Lp); n == 0 blends with the conditioning utterance and, when ``c`` is given, removes the gain."""
        B, Lp = audio.shape
        Tf = Lp // dsp.HOP + 1
        eng = self.dft_engine
        apad, _ = ops.rms_pad(audio.contiguous(), Lp, normalize=False)
        in3 = dsp.stft_in3(apad, Tf, eng)
        st = torch.tensor([float(T[n])], dtype=torch.float32, device=audio.device)      # :252: a float32 array entry -> interpolated embedding
        est = self.model.forward_in3(in3, cond_in3, st)
        z = torch.empty(B * Tf, dsp.LDZ, device=audio.device, dtype=torch.float32)
        ops.decompress_rows(est, z)
        pred = dsp.istft_rows(z, B, Tf, None, eng)
        if n > 0:
            if noise is None:
                raise RuntimeError("reverse step n > 0 needs the Gaussian draw")
            return ops.diffusion_update(audio, noisy_audio, pred, noise.to(torch.float32).contiguous(), float(c1[n]), float(c2[n]), -float(c3[n]),
                                        float(delta_bar[n]) ** 0.5)
        g = GAMMA0
        return ops.diffusion_update(audio, noisy_audio, pred, None, (1.0 - g) * float(c1[n]), g, -(1.0 - g) * float(c3[n]), 0.0, c_div=c)


@torch.no_grad()
def predict_tsc(model, args, config, noisy_signal, alpha, beta, alpha_cum, sigmas, T, c1, c2, c3, delta, delta_bar,
                device=torch.device("cuda"), noise_fn=None):
    """inference_diffuse.predict_tsc: 1-D numpy noisy utterance -> 1-D numpy enhanced utterance of the same length."""
    if getattr(args, "comp_type", "pow") != "pow" or config.N_FFT != dsp.N_FFT or config.HOP_SAMPLES != dsp.HOP:
        raise RuntimeError("se_b200 DSP kernels are specialised for n_fft=400, hop=100, comp_type='pow'")
    if len(alpha) != len(c1):
        raise RuntimeError("alpha and c1 must have one entry per inference step")
    x = torch.from_numpy(np.ascontiguousarray(noisy_signal, dtype=np.float32)).to(device)
    y = DiffusionEnhancerB200(model).reverse(x.unsqueeze(0), T, c1, c2, c3, delta_bar, noise_fn)
    return torch.flatten(y).cpu().numpy()
