"""GPU: backward of the DSP bracket -- compressed_stft / uncompressed_istft as autograd Functions with
hand-written backward kernels -- against torch autograd through the oracle's float64 evaluation of the same functions, on the
consistency-loss chain of train_gan (core/function.py:227-254) at the training configuration's shape (batch 4, 2 s crops)."""
import pytest
import torch

from conftest import rel_l2, rel_max
from oracle import tscnet_oracle as O, weights

import se_b200

pytestmark = pytest.mark.gpu
DEV = "cuda"
GRAD_TOL = 1e-3            # of the largest gradient entry; the forward tolerances of the path are 1e-3 / 1e-4 as well


def _w(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


def test_compressed_stft_backward_matches_autograd():
    noisy, _ = weights.synth_wave(4, 32000, 5, "speech")
    x_o = (3.0 * noisy).double().requires_grad_(True)
    spec_o = O.compressed_stft(x_o, window=O.hamming_periodic().double())
    wr, wi = _w(spec_o.shape, 1), _w(spec_o.shape, 2)
    loss_o = (spec_o.real * wr).sum() + (spec_o.imag * wi).sum() + spec_o.abs().pow(2).sum()
    loss_o.backward()
    x_g = (3.0 * noisy).to(DEV).requires_grad_(True)
    spec_g = se_b200.compressed_stft(x_g)
    assert spec_g.requires_grad and spec_g.shape == (4, 201, 321)
    loss_g = (spec_g.real * wr.float().to(DEV)).sum() + (spec_g.imag * wi.float().to(DEV)).sum() + spec_g.abs().pow(2).sum()
    loss_g.backward()
    assert rel_max(torch.view_as_real(spec_g.detach().cpu()), torch.view_as_real(spec_o.detach())) < 2e-4
    err = rel_max(x_g.grad.cpu(), x_o.grad)
    assert err < GRAD_TOL, f"d loss / d waveform: {err:.3e}"
    # no graph, no autograd Function: the inference path is untouched
    with torch.no_grad():
        assert not se_b200.compressed_stft(x_g).requires_grad


def test_stft_gradient_at_full_size_matches_fp32_autograd():
    """BASELINE-size check (64 x 4 s).  The compression's Jacobian ~ |X|^-0.7 is unbounded on empty bins (speech-like input has bins
    down to 4e-8 of peak), where ANY fp32 evaluation differs from float64 by O(1); the loss therefore reads the bins above 5 % of the
    compressed peak only.  There the default training path must be as close to float64 autograd as torch's own fp32 autograd on the
    reference's code (torch.stft on the GPU): the hand-written backward adds no error of its own."""
    noisy, _ = weights.synth_wave(64, 64000, 7, "speech")
    x = (3.0 * noisy).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(1)
    w1, w2 = torch.randn(64, 201, 641, device=DEV, generator=gen), torch.randn(64, 201, 641, device=DEV, generator=gen)
    win = torch.hamming_window(400, device=DEV)

    def ref_stft(sig, window):      # core/function.py:685-693 on the GPU (cuFFT): the reference's own code path
        s = torch.stft(sig, 400, 100, window=window, onesided=True, return_complex=True)
        mag, ph = s.abs() ** 0.3, s.angle()
        return torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))

    with torch.no_grad():
        mag = ref_stft(x.double(), win.double()).abs()
        mask = mag > 0.05 * mag.max()
    assert 0.5 < float(mask.float().mean()) < 1.0

    def grad(stft, xin, dt):
        xi = xin.clone().requires_grad_(True)
        s = stft(xi)
        ((s.real * (w1 * mask).to(dt)).sum() + (s.imag * (w2 * mask).to(dt)).sum()).backward()
        return xi.grad

    g64 = grad(lambda a: ref_stft(a, win.double()), x.double(), torch.float64)
    e_ours = rel_l2(grad(se_b200.compressed_stft, x, torch.float32), g64)
    e_t32 = rel_l2(grad(lambda a: ref_stft(a, win), x, torch.float32), g64)
    assert e_ours < max(3.0 * e_t32, 1e-3), f"ours {e_ours:.3e} vs torch fp32 {e_t32:.3e} (both against float64)"
    print(f"full-size STFT gradient rel-L2 vs float64: ours {e_ours:.3e}, torch fp32 {e_t32:.3e}")


def test_uncompressed_istft_backward_matches_autograd():
    g = torch.Generator().manual_seed(3)
    y_re, y_im = torch.randn(4, 201, 81, generator=g), torch.randn(4, 201, 81, generator=g)
    spec_o = torch.complex(y_re.double(), y_im.double()).requires_grad_(True)
    wav_o = O.uncompressed_istft(spec_o, window=O.hamming_periodic().double())
    w = _w(wav_o.shape, 4)
    loss_o = (wav_o * w).sum() + wav_o.abs().sum()
    loss_o.backward()
    spec_g = torch.complex(y_re, y_im).to(DEV).requires_grad_(True)
    wav_g = se_b200.uncompressed_istft(spec_g)
    loss_g = (wav_g * w.float().to(DEV)).sum() + wav_g.abs().sum()
    loss_g.backward()
    assert rel_max(wav_g.detach().cpu(), wav_o.detach()) < 1e-4
    err = rel_max(torch.view_as_real(spec_g.grad.cpu()), torch.view_as_real(spec_o.grad))
    assert err < GRAD_TOL, f"d loss / d spectrogram: {err:.3e}"


@pytest.mark.parametrize("dft_engine", ["tcgen05", None])
def test_consistency_loss_chain_gradient(dft_engine):
    """train_gan's 'scp' branch (core/function.py:227-254): est_complex -> uncompressed_istft -> est_audio -> compressed_stft ->
    magnitude / real-imaginary MSE against the clean* pipeline + L1 time loss; gradient w.r.t. the generator's output.

    The compression's Jacobian scales with |X|^-0.7, so near-empty bins amplify the forward DFT's absolute error floor: measured
    against the float64 oracle, torch's own fp32 autograd is 3.5e-4 (max) / 3.9e-5 (rel-L2) off on this chain, the fp32 FFMA DFT
    engine -- the default under autograd (dsp.GRAD_DFT_ENGINE, engine=None here) -- matches that, and the split-bf16 tensor-core DFT
    (three planes, ~1e-6 of peak absolute) lands at 1.7e-3 / 1.9e-4 with the worst entry in bin 199 of 201.  Bounds: 1e-3 max for the
    default; 5e-3 max and 1e-3 rel-L2 for the tensor-core engine."""
    noisy, clean = weights.synth_wave(4, 32000, 7, "speech")
    with torch.no_grad():
        c = torch.sqrt(noisy.shape[-1] / torch.sum(noisy ** 2.0, dim=-1, keepdim=True))
        clean_spec = O.compressed_stft(clean * c)
        est0 = O.compressed_stft(noisy * c)            # stands in for the generator's output (same statistics)

    def chain(est, cspec, istft, stft, mse):
        est_audio = istft(est)
        est_prime = stft(est_audio)
        clean_prime_audio = istft(cspec)
        clean_prime = stft(clean_prime_audio)
        loss_mag = mse(est_prime.abs(), clean_prime.abs())
        loss_ri = mse(est_prime.real, clean_prime.real) + mse(est_prime.imag, clean_prime.imag)
        time_loss = torch.mean(torch.abs(est_audio - clean_prime_audio))
        return 0.9 * loss_mag + 0.1 * loss_ri + 0.2 * time_loss, est_audio

    mse = torch.nn.functional.mse_loss
    w64 = O.hamming_periodic().double()
    est_o = est0.to(torch.complex128).requires_grad_(True)
    loss_o, audio_o = chain(est_o, clean_spec.to(torch.complex128), lambda s: O.uncompressed_istft(s, window=w64),
                            lambda a: O.compressed_stft(a, window=w64), mse)
    loss_o.backward()
    est_g = est0.to(DEV).requires_grad_(True)
    loss_g, audio_g = chain(est_g, clean_spec.to(DEV), lambda s: se_b200.uncompressed_istft(s, engine=dft_engine),
                            lambda a: se_b200.compressed_stft(a, engine=dft_engine), mse)
    loss_g.backward()
    assert abs(float(loss_g.detach()) - float(loss_o.detach())) / abs(float(loss_o.detach())) < 1e-3
    assert rel_max(audio_g.detach().cpu(), audio_o.detach()) < 1e-4
    gg, go = torch.view_as_real(est_g.grad.cpu()), torch.view_as_real(est_o.grad)
    err, err2 = rel_max(gg, go), rel_l2(gg, go)
    assert err < (GRAD_TOL if dft_engine is None else 5 * GRAD_TOL) and err2 < GRAD_TOL, f"d loss / d est_complex: max {err:.3e}, rel-L2 {err2:.3e}"
