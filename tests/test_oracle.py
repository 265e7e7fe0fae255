"""CPU: the oracle restatement against the committed golden vectors (produced by the unmodified reference,
oracle/make_golden.py) and -- when /root/reference is mounted -- against the live reference stage by stage."""
import numpy as np
import pytest
import torch

from conftest import rel_max, reverse_noises
from oracle import ref_import, tscnet_oracle as O, weights

CASES = ["speech_b2_L8000", "noise_b1_L4050_wrap"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(golden, name):
    g = golden(name)
    sd = weights.synth_state_dict(int(g["weight_seed"]))
    noisy = torch.from_numpy(g["noisy"])
    st = {}
    with torch.no_grad():
        y = O.predict(noisy, sd, stages=st)
    spec_ref = torch.complex(torch.from_numpy(g["spec_real"]), torch.from_numpy(g["spec_imag"]))
    assert rel_max(torch.view_as_real(st["spec"]), torch.view_as_real(spec_ref)) < 2e-5
    with torch.no_grad():
        fr, fi = O.tscnet_forward(spec_ref, sd)
    assert rel_max(fr, torch.from_numpy(g["final_real"])) < 1e-5
    assert rel_max(fi, torch.from_numpy(g["final_imag"])) < 1e-5
    assert rel_max(y, torch.from_numpy(g["enhanced"])) < 2e-5
    assert y.shape == noisy.shape


DIFF_STEPS = [("int1", [7], torch.int64), ("frac1", [3.4], torch.float32), ("intB", [2, 40], torch.int64)]


def test_diffusion_oracle_matches_golden(golden):
    """the tsc_diffusion.TSCNet restatement against outputs of the reference's own module."""
    g = golden("diffusion_b2_L3000")
    sd = weights.synth_state_dict(int(g["weight_seed"]), spec=weights.tsc_diffusion_spec())
    sx = torch.complex(torch.from_numpy(g["spec_x_real"]), torch.from_numpy(g["spec_x_imag"]))
    sn = torch.complex(torch.from_numpy(g["spec_n_real"]), torch.from_numpy(g["spec_n_imag"]))
    for tag, vals, dt in DIFF_STEPS:
        with torch.no_grad():
            fr, fi = O.tsc_diffusion_forward(sx, sn, torch.tensor(vals, dtype=dt), sd, int(g["max_steps"]))
        assert rel_max(fr, torch.from_numpy(g[f"final_real_{tag}"])) < 1e-5, tag
        assert rel_max(fi, torch.from_numpy(g[f"final_imag_{tag}"])) < 1e-5, tag
    # the step matters (different steps give different outputs) and an integral float step equals the integer one
    assert rel_max(torch.from_numpy(g["final_real_int1"]), torch.from_numpy(g["final_real_frac1"])) > 1e-3
    with torch.no_grad():
        a = O.diffusion_embedding(torch.tensor([7]), sd, "merge_block.diffusion_embedding", 50)
        b = O.diffusion_embedding(torch.tensor([7.0]), sd, "merge_block.diffusion_embedding", 50)
    assert torch.equal(a, b)


def test_predict_tsc_oracle_matches_golden(golden):
    """the reverse process (inference_diffuse.predict_tsc, 6-step fast schedule, wrap-pad branch) restated vs the reference's output"""
    g = golden("diffusion_reverse_L2950")
    sd = weights.synth_state_dict(int(g["weight_seed"]), spec=weights.tsc_diffusion_spec())
    with torch.no_grad():
        y = O.predict_tsc(torch.from_numpy(g["noisy"]), sd, int(g["max_steps"]), g["T"], g["c1"], g["c2"], g["c3"], g["delta_bar"],
                          reverse_noises(g))
    assert y.shape == g["noisy"].shape
    assert rel_max(y, torch.from_numpy(g["enhanced"])) < 5e-5


def test_diffusion_spec_state_dict_contract():
    spec = weights.tsc_diffusion_spec()
    assert len(spec) == 401 and len({k for k, _, _ in spec}) == 401
    assert sum(k.startswith("dense_encoder_noisy.") for k, _, _ in spec) == 30
    assert sum(k.startswith("merge_block.") for k, _, _ in spec) == 12


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")
def test_diffusion_oracle_against_live_reference():
    ref = ref_import.load()
    spec = weights.tsc_diffusion_spec()
    model = ref.DiffusionTSCNet(num_channel=64, num_features=201, noise_schedule=list(range(30)))
    assert list(model.state_dict().keys()) == [k for k, _, _ in spec]
    sd = weights.synth_state_dict(5, spec=spec)
    model.load_state_dict(sd, strict=True)
    model.eval()
    est, cond = weights.synth_wave(1, 2400, 3, "speech")
    with torch.no_grad():
        sx, sn = O.compressed_stft(est), O.compressed_stft(cond)
        for step in (torch.tensor([29]), torch.tensor([0.25])):
            fr, fi = model(sx, sn, step)
            ofr, ofi = O.tsc_diffusion_forward(sx, sn, step, sd, 30)
            assert rel_max(ofr, fr) < 1e-5 and rel_max(ofi, fi) < 1e-5


def test_oracle_matches_golden_default_init(golden):
    """second weight set: PyTorch-default initialisation, output of the reference's predict()"""
    g = golden("default_init_b1_L6000")
    sd = weights.torch_default_state_dict(int(g["weight_seed"]))
    assert list(sd.keys()) == [k for k, _, _ in weights.tscnet_spec()]
    with torch.no_grad():
        y = O.predict(torch.from_numpy(g["noisy"]), sd)
    assert rel_max(y, torch.from_numpy(g["enhanced"])) < 2e-5


def test_golden_inputs_are_reproducible(golden):
    """the committed inputs are exactly what oracle.weights regenerates from the stored seeds"""
    g = golden("speech_b2_L8000")
    noisy, clean = weights.synth_wave(2, 8000, int(g["wave_seed"]), "speech")
    assert np.array_equal(noisy.numpy(), g["noisy"]) and np.array_equal(clean.numpy(), g["clean"])


def test_spec_state_dict_contract():
    spec = weights.tscnet_spec()
    assert len(spec) == 359
    sd = weights.synth_state_dict(3)
    n_params = sum(v.numel() for k, v in sd.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_params == 1_834_833


def test_istft_inverts_stft():
    x = 0.1 * torch.randn(2, 3200, generator=torch.Generator().manual_seed(0))
    w = O.hamming_periodic()
    fr = O.stft_frames(x) * w
    spec = torch.fft.rfft(fr, dim=-1).transpose(1, 2)
    y = O.uncompressed_istft(O.power_compress(spec))        # decompress(compress(S)) == S
    assert rel_max(y, x) < 1e-4


def test_attention_chunking_is_exact():
    sd = weights.synth_state_dict(0)
    x = torch.randn(6, 37, 64, generator=torch.Generator().manual_seed(1))
    p = "TSCB_1.time_conformer.attn"
    with torch.no_grad():
        a, b = O.attention(x, sd, p), O.attention(x, sd, p, chunk=4)
    assert torch.equal(a, b)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")
def test_oracle_against_live_reference():
    ref = ref_import.load()
    torch.manual_seed(0)
    model = ref.TSCNet(num_channel=64, num_features=201)
    model.apply(ref.kaiming_init)            # main_gan.py:147
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    noisy, _ = weights.synth_wave(1, 6000, 9, "speech")
    caught = {}
    hooks = [model.dense_encoder.register_forward_hook(lambda m, i, o: caught.__setitem__("encoder", o)),
             model.TSCB_2.register_forward_hook(lambda m, i, o: caught.__setitem__("tscb2", o)),
             model.mask_decoder.register_forward_hook(lambda m, i, o: caught.__setitem__("mask", o)),
             model.complex_decoder.register_forward_hook(lambda m, i, o: caught.__setitem__("complex", o))]
    y_ref = ref.predict(model, ref.config, noisy[0].numpy(), device=torch.device("cpu"))
    for h in hooks:
        h.remove()
    st = {}
    with torch.no_grad():
        y = O.predict(noisy, sd, stages=st)
    for k in ("encoder", "tscb2", "mask", "complex"):
        assert rel_max(st[k], caught[k]) < 1e-5, k
    assert rel_max(y[0], torch.from_numpy(y_ref)) < 2e-5
    # long-sequence path with the +-512 clamp active, attention evaluated in chunks
    x = torch.randn(3, 600, 64, generator=torch.Generator().manual_seed(2))
    blk = model.TSCB_1.time_conformer
    with torch.no_grad():
        a = blk(x)
        b = O.conformer_block(x, sd, "TSCB_1.time_conformer", chunk=2)
    assert rel_max(b, a) < 1e-5


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")
def test_batch_stft_against_live_reference():
    """normalize_batch / batch_stft (core/function.py:647-683) restated vs the reference itself."""
    import types
    ref = ref_import.load()
    noisy, clean = weights.synth_wave(3, 3200, 21, "speech")
    out = ref.batch_stft({"audio": clean.clone(), "noisy": noisy.clone()}, types.SimpleNamespace(gpu=None), ref.config)
    r_clean, r_noisy, r_cspec, r_nspec, r_creal, r_cimag, ones, win = out
    o_clean, o_noisy, o_cspec, o_nspec = O.batch_stft(clean, noisy)
    assert rel_max(o_clean, r_clean) < 1e-6 and rel_max(o_noisy, r_noisy) < 1e-6
    assert rel_max(torch.view_as_real(o_cspec), torch.view_as_real(r_cspec)) < 2e-5
    assert rel_max(torch.view_as_real(o_nspec), torch.view_as_real(r_nspec)) < 2e-5
    assert r_creal.shape == (3, 1, 201, 33) and ones.shape == (3,) and win.shape == (400,)
    # the noisy signal ends up with unit RMS; the clean one shares its gain
    assert torch.allclose(r_noisy.pow(2).mean(-1), torch.ones(3), atol=1e-5)


# ---- train mode: the restatement with injected dropout masks / BatchNorm batch statistics, and its autograd
# ---- gradients, against the unmodified reference under .train() (tests/golden/train_b2_L10000.npz, oracle/make_golden.py --train).
# ---- The golden's gradients come from the reference run in float64; `ref32_err` is the reference's own float32 distance from them.
def train_oracle_run(g, dtype=torch.float32, device="cpu"):
    """(final_real, final_imag, running-stat dict, {param: grad}) of the oracle for the training golden's inputs"""
    import synth
    sd = {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device)) for k, v in synth.synth_state_dict(int(g["weight_seed"])).items()}
    params = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    for k in params:
        sd[k].requires_grad_(True)
    spec = torch.complex(torch.from_numpy(g["spec_real"]), torch.from_numpy(g["spec_imag"])).to(device)
    if dtype == torch.float64:
        spec = spec.to(torch.complex128)
    B, _, T = spec.shape
    masks = {k: v.to(device) for k, v in synth.masks_reference_layout(synth.dropout_masks(int(g["mask_seed"]), B, T, 101)).items()}
    tr = O.TrainCtx(masks)
    fr, fi = O.tscnet_forward(spec, sd, tr=tr)
    gr, gi = synth.cotangents(int(g["cot_seed"]), B, T)
    ((fr * gr.to(fr)).sum() + (fi * gi.to(fi)).sum()).backward()
    return fr.detach(), fi.detach(), tr.running, {k: sd[k].grad for k in params}


def grad_errors(grads, g):
    """per-parameter rel-L2 of `grads` against the golden's float64 gradients; zero-gradient biases are bounded in magnitude instead"""
    import synth
    gmax = max(float(np.abs(g["grad:" + k]).max()) for k in grads)
    errs = {}
    for k, gv in grads.items():
        ref = torch.from_numpy(g["grad:" + k]).double()
        if synth.has_zero_gradient(k):
            assert float(gv.abs().max()) < 1e-3 * gmax, k
            continue
        errs[k] = float((gv.detach().double().cpu() - ref).norm() / ref.norm().clamp_min(1e-30))
    return errs


def test_oracle_train_mode_matches_reference_golden(golden):
    g = golden("train_b2_L10000")
    peak = np.abs(g["final_real"]).max()
    # float32 restatement: forward and BatchNorm buffers against the reference's float32 run
    fr, fi, running, grads32 = train_oracle_run(g, torch.float32)
    assert (fr - torch.from_numpy(g["final_real"])).abs().max() / peak < 2e-5
    assert (fi - torch.from_numpy(g["final_imag"])).abs().max() / peak < 2e-5
    assert len(running) == 24
    for k, v in running.items():
        ref = torch.from_numpy(np.asarray(g["buf:" + k]))
        assert torch.allclose(v.to(ref.dtype), ref, rtol=1e-5, atol=1e-6), k
    # float64 restatement: the same function as the reference in float64 -> gradients agree to float32 storage rounding
    _, _, _, grads64 = train_oracle_run(g, torch.float64)
    assert len(grads64) == 335
    e64 = grad_errors(grads64, g)
    worst = max(e64, key=e64.get)
    assert e64[worst] < 1e-6, (worst, e64[worst])
    # float32 restatement: no further from the float64 truth than the reference's own float32 run (x3 slack, floor 1e-4)
    e32 = grad_errors(grads32, g)
    for k, e in e32.items():
        assert e < max(1e-4, 3.0 * float(g["ref32_err:" + k])), (k, e, float(g["ref32_err:" + k]))
