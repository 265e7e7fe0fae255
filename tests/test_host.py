"""CPU: host-side logic -- state-dict contract, weight packing, the C-ABI surface, error behaviour, sharding."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_max
from oracle import tscnet_oracle as O, weights

import se_b200
from se_b200 import packing


def test_state_dict_is_the_reference_contract():
    m = se_b200.TSCNet(num_channel=64, num_features=201)
    sd = weights.synth_state_dict(0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape and v.dtype == sd[k].dtype, k
    m.load_state_dict(sd, strict=True)
    # checkpoints saved from DataParallel/DDP carry 'module.' (inference_gan.py:66-68)
    bn = m.TSCB_1.time_conformer.conv.net[5]
    assert isinstance(bn, torch.nn.BatchNorm1d)
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
    assert sum(isinstance(x, torch.nn.SyncBatchNorm) for x in conv.modules()) == 8


def test_diffusion_state_dict_is_the_reference_contract():
    """tsc_diffusion.TSCNet(num_channel, num_features, noise_schedule) -- same keys, order, shapes"""
    from se_b200 import tsc_diffusion
    m = tsc_diffusion.TSCNet(num_channel=64, num_features=201, noise_schedule=[0.1] * 50)
    sd = weights.synth_state_dict(0, spec=weights.tsc_diffusion_spec())
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape and v.dtype == sd[k].dtype, k
    m.load_state_dict(sd, strict=True)
    tab = m.merge_block.diffusion_embedding.embedding          # non-persistent buffer, as models/DiffuSE.py:42
    assert tab.shape == (50, 128) and torch.equal(tab, O.diffusion_step_table(50))
    with pytest.raises(TypeError):
        tsc_diffusion.TSCNet(64, 201)
    m.eval()
    z = torch.zeros(1, 201, 5, dtype=torch.complex64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(z, z, torch.tensor([3]))
    with pytest.raises(TypeError):
        m(z)


def test_no_cpu_path():
    m = se_b200.TSCNet().eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 201, 5, dtype=torch.complex64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        se_b200.EnhancerB200(m)(torch.zeros(1, 800))
    with pytest.raises(RuntimeError):
        se_b200.compressed_stft(torch.zeros(1, 800))


def test_swizzle_is_an_involution_and_matches_formula():
    blk = torch.arange(16 * 64, dtype=torch.float32).reshape(16, 64).to(torch.bfloat16)
    sw = packing.swizzle128(blk)
    assert torch.equal(packing.swizzle128(sw), blk)
    raw = sw.view(torch.int16).reshape(-1)
    src = blk.view(torch.int16)
    for r in (0, 3, 9, 15):
        for c in range(8):
            off = (r * 128 + ((c ^ (r & 7)) << 4)) // 2
            assert torch.equal(raw[off:off + 8], src[r, c * 8:(c + 1) * 8])


def _unpack_tc(pw):
    img = pw.w_tc.view(torch.bfloat16).reshape(pw.tc_ntiles, pw.K // 64, 2, pw.tc_ntile, 64)
    W = torch.zeros(pw.tc_ntiles * pw.tc_ntile, pw.K)
    for j in range(pw.tc_ntiles):
        for kc in range(pw.K // 64):
            hi = packing.swizzle128(img[j, kc, 0]).float()
            lo = packing.swizzle128(img[j, kc, 1]).float()
            W[j * pw.tc_ntile:(j + 1) * pw.tc_ntile, kc * 64:(kc + 1) * 64] = hi + lo
    return W


def test_pack_weight_images():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(402, 400, generator=g)
    pw = packing.pack_weight(w, 208)
    assert (pw.K, pw.tc_ntiles, pw.simt_npad) == (448, 2, 448)
    W = _unpack_tc(pw)
    assert rel_max(W[:402, :400], w) < 2.0 ** -16       # hi + lo keeps ~17 mantissa bits
    assert W[402:].abs().max() == 0 and W[:, 400:].abs().max() == 0
    assert torch.equal(pw.w_simt[:400, :402], w.t())


def test_split_bf16_three_product_error_model():
    """CPU emulation of the tensor path's arithmetic (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulate)."""
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(256, 384, generator=g), torch.randn(64, 384, generator=g) * 0.07
    ah, al = packing.split_bf16(a)
    bh, bl = packing.split_bf16(b)
    f = lambda x, y: x.float() @ y.float().t()
    approx = f(ah, bh) + f(ah, bl) + f(al, bh)
    exact = (a.double() @ b.double().t()).float()
    assert rel_max(approx, exact) < 3e-5
    assert rel_max(f(ah, bh), exact) > 1e-3             # single-pass bf16 is not enough


def test_three_plane_split_is_fp32_grade():
    """six products of (hi, mid, lo) planes: the DFT mode of the tensor path reproduces an fp32 GEMM"""
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(128, 448, generator=g), torch.randn(96, 448, generator=g)
    ah, am, al = packing.split_bf16_3(a)
    bh, bm, bl = packing.split_bf16_3(b)
    assert rel_max(ah.float() + am.float() + al.float(), a) < 2.0 ** -22
    f = lambda x, y: x.float() @ y.float().t()
    approx = f(al, bh) + f(ah, bl) + f(am, bm) + f(am, bh) + f(ah, bm) + f(ah, bh)
    exact = (a.double() @ b.double().t()).float()
    assert rel_max(approx, exact) < 2e-6
    pw = packing.pack_weight(b, 48, planes=3)
    assert pw.planes == 3 and pw.w_tc.numel() == 2 * 7 * 3 * 48 * 64 * 2


def test_dft_bases_match_torch_fft():
    x = torch.randn(3, 400, generator=torch.Generator().manual_seed(2))
    w = O.hamming_periodic()
    S = torch.fft.rfft(x * w, dim=-1)
    got = (x.double() @ packing.dft_basis().double().t()).reshape(3, 201, 2)
    assert rel_max(got[..., 0], S.real) < 1e-5 and rel_max(got[..., 1], S.imag) < 1e-5
    Z = torch.randn(3, 201, dtype=torch.complex64, generator=torch.Generator().manual_seed(3))
    fr = torch.fft.irfft(Z, n=400, dim=-1) * w
    z = torch.view_as_real(Z).reshape(3, 402)
    got = z.double() @ packing.idft_basis().double().t()
    assert rel_max(got, fr) < 1e-5
    env = packing.inv_envelope(9)
    assert env.shape == (800,) and float(env.max()) < 1.0 and float(env.min()) > 0.5


def test_abi_exports_every_declared_symbol():
    lib_path = se_b200._lib.lib_path()
    if not os.path.exists(lib_path):
        se_b200._lib.load()            # builds with nvcc (cross-compiles without a GPU)
    hdr = open(os.path.join(ROOT, "include", "seb200.h")).read()
    declared = sorted(set(re.findall(r"\b(seb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == se_b200._lib.EXPORTS
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.seb200_version() == se_b200._lib.ABI_VERSION


@pytest.mark.parametrize("N,K,ntile,planes", [(64, 64, 64, 2), (192, 64, 192, 2), (64, 1536, 64, 2), (402, 400, 208, 3), (128, 128, 128, 2), (70, 100, 64, 2)])
def test_c_abi_weight_packer_matches_torch_restatement(N, K, ntile, planes):
    """seb200_pack_weights (host-side C, what a non-Python caller uses) == the torch restatement of the layout, bit for bit"""
    w = torch.randn(N, K, generator=torch.Generator().manual_seed(N + K)) * (K ** -0.5)
    w[0, 0], w[-1, -1] = 1e-30, 3.0e38                 # a denormal-range value and a near-max value go through the rounding too
    a, b = packing.pack_weight(w, ntile, None, planes), packing.pack_weight_torch(w, ntile, None, planes)
    assert (a.N, a.K, a.tc_ntile, a.tc_ntiles, a.simt_npad, a.planes) == (b.N, b.K, b.tc_ntile, b.tc_ntiles, b.simt_npad, b.planes)
    assert torch.equal(a.w_tc, b.w_tc) and torch.equal(a.w_simt, b.w_simt)


def test_workspace_bytes_matches_the_host_allocation():
    """seb200_workspace_bytes(kind, B, T, F) == what TSCNet.workspace really allocates"""
    from se_b200 import tsc_diffusion
    lib = se_b200._lib.load()
    B, T, F = 2, 7, 201

    def total(ws):
        n = 0
        for v in ws.values():
            for t in (v if isinstance(v, list) else [v]):
                n += t.numel() * t.element_size()
        return n
    assert lib.seb200_workspace_bytes(0, B, T, F) == total(se_b200.TSCNet().workspace(B, T, "cpu"))
    assert lib.seb200_workspace_bytes(1, B, T, F) == total(tsc_diffusion.TSCNet(64, 201, [0.1] * 4).workspace(B, T, "cpu"))
    assert lib.seb200_workspace_bytes(0, 64, 641, 201) < 45e9            # configs[1]: ~ 37 GB of the 180 GB
    assert lib.seb200_workspace_bytes(0, 1, 5, 200) == -1 and b"workspace_bytes" in lib.seb200_last_error_string()


def test_abi_rejects_bad_arguments_without_a_gpu():
    lib = se_b200._lib.load()
    assert lib.seb200_gemm(None, 0, None) == -1
    assert b"null descriptor" in lib.seb200_last_error_string()
    g = se_b200._lib.SebGemm()
    g.M, g.N, g.K = 128, 64, 100                       # K not a multiple of 64
    assert lib.seb200_gemm(ctypes.byref(g), 0, None) == -1
    assert lib.seb200_rms_pad(None, 1, 100, 100, 1, None, None, None) == -1


def test_abi_validates_merge_block_descriptors_without_a_gpu():
    """argument checks of the two-source loader / gate epilogue happen before any launch"""
    lib = se_b200._lib.load()
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf) & ~15
    g = se_b200._lib.SebGemm()
    g.loader, g.epilogue = se_b200._lib.LOAD_ROWS2, se_b200._lib.EPI_GATE
    g.M, g.N, g.K, g.lda, g.ldo = 128, 128, 64, 64, 64          # K must be 128
    g.a[0], g.a[1], g.out = addr, addr, addr
    assert lib.seb200_gemm(ctypes.byref(g), 0, None) == -2 and b"two-source" in lib.seb200_last_error_string()
    g.K, g.a[1] = 128, 0                                          # second source missing
    assert lib.seb200_gemm(ctypes.byref(g), 0, None) == -2
    g.a[1], g.resid, g.ldr = addr, addr, 0                        # row bias without rows-per-group
    assert lib.seb200_gemm(ctypes.byref(g), 0, None) == -1 and b"rows per group" in lib.seb200_last_error_string()
    assert lib.seb200_diffusion_embed(None, 1, None, 50, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.seb200_stft_fold(None, 1, 5, 400, 400, None, None) == -1
    assert lib.seb200_pack_weights(None, 64, 64, 64, 2, None, None) == -1


def test_c_abi_weight_packer_random_shapes():
    """hypothesis: any (N, K, n-tile, planes) the engine accepts packs identically in C and in the torch restatement"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 300), st.integers(1, 300), st.sampled_from([16, 64, 128, 192, 208, 256]), st.sampled_from([2, 3]), st.integers(0, 1000))
    def check(N, K, ntile, planes, seed):
        w = torch.randn(N, K, generator=torch.Generator().manual_seed(seed))
        a, b = packing.pack_weight(w, ntile, None, planes), packing.pack_weight_torch(w, ntile, None, planes)
        assert torch.equal(a.w_tc, b.w_tc) and torch.equal(a.w_simt, b.w_simt) and (a.K, a.tc_ntiles, a.simt_npad) == (b.K, b.tc_ntiles, b.simt_npad)
    check()


def test_shard_slice_partitions():
    for n, g in [(4096, 8), (10, 4), (3, 8), (64, 1)]:
        idx = []
        for r in range(g):
            s = se_b200.shard_slice(n, r, g)
            idx += list(range(s.start, s.stop))
        assert idx == list(range(n))


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import se_b200
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
waves = torch.arange(7 * 5, dtype=torch.float32).reshape(7, 5)
fake = lambda x: x * 2.0 + 1.0                      # stands in for the per-rank CUDA enhancer
out = se_b200.enhance_sharded(fake, waves, rank, world, micro_batch=2, gather=True)
assert torch.equal(out, fake(waves)), (rank, out)
local = se_b200.enhance_sharded(fake, waves, rank, world, micro_batch=3)
sl = se_b200.shard_slice(7, rank, world)
assert torch.equal(local, fake(waves[sl]))
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()


# ---- training: the data-parallel exchange (SURVEY 8e training; main_gan.py:168-171): one flat all-reduce of the gradient buffer ---------
_GLOO_TRAIN_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import se_b200
from se_b200 import training
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
m = se_b200.TSCNet(64, 201)
st = training._state(m)
flat, views = st.flat(0, torch.device("cpu"))
names, params = st.params()
assert flat.numel() == 1834833 and len(names) == 335
# what GeneratorFunction.backward hands to autograd: views of the flat buffer become the .grad tensors
flat.copy_(torch.arange(flat.numel(), dtype=torch.float32) * (rank + 1))
for n, p in zip(names, params):
    p.grad = views[n].view(views[n].shape)
assert st.grad_buffer() is flat
# the next backward must not write the buffer the live .grad tensors alias (it would double-count under accumulation)
other, _ = st.pick_flat(torch.device("cpu"))
assert other is not flat and other.data_ptr() != flat.data_ptr()
out = se_b200.allreduce_gradients(m)
assert out is flat
expect = torch.arange(flat.numel(), dtype=torch.float32) * (1 + 2) / 2.0           # mean over the two ranks
assert torch.allclose(flat, expect)
assert torch.equal(params[5].grad.reshape(-1), flat[sum(p.numel() for p in params[:5]):sum(p.numel() for p in params[:6])])
# gradients that do not alias the flat buffer (e.g. set by another code path) still get averaged (coalesced fallback)
for p in params:
    p.grad = torch.full_like(p, float(rank))
assert st.grad_buffer() is None
se_b200.allreduce_gradients(m)
assert all(torch.allclose(p.grad, torch.full_like(p, 0.5)) for p in params)
# SyncBatchNorm conversion (main_gan.py:154) keeps the parameter tree: eight SyncBatchNorm children where the BatchNorm1d were
ms = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
assert sum(isinstance(x, torch.nn.SyncBatchNorm) for x in ms.modules()) == 8
assert list(ms.state_dict().keys()) == list(se_b200.TSCNet(64, 201).state_dict().keys())
grp, w = training._sync_group(ms, training._state(ms))
assert grp is not None and w == 2
assert training._sync_group(se_b200.TSCNet(64, 201), st) == (None, 1)               # plain BatchNorm1d: local statistics
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "t.py"
    script.write_text(_GLOO_TRAIN_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out.decode()


def test_training_dropout_sites_match_the_test_generator():
    import synth
    from se_b200 import training
    assert training.dropout_sites() == synth.dropout_sites() and len(training.dropout_sites()) == 40


# ------------------------------------------------------------------------------------------------ PESQ label pipeline (host side)
def _toy_score(sr, c, n):
    """stand-in scorer with PESQ's range (1 .. 4.5), monotone in the correlation of the pair; raises on a silent reference like the real one"""
    if float(np.abs(c).max()) == 0.0:
        raise RuntimeError("silent")
    r = float(np.dot(c, n) / (np.linalg.norm(c) * np.linalg.norm(n) + 1e-12))
    return 1.0 + 3.5 * max(r, 0.0)


def test_metric_label_pipeline_matches_reference_rule():
    """batch_pesq (models/discriminator.py:17-32): label = (score - 1) / 3.5, a raising scorer counts as -1, order preserved, asynchronous submit"""
    import time
    from se_b200 import MetricLabelPipeline, batch_pesq
    g = torch.Generator().manual_seed(0)
    clean = torch.randn(6, 4000, generator=g)
    est = clean[:, :3900] + 0.5 * torch.randn(6, 3900, generator=g)          # shorter than clean: cropped like core/function.py:283-285
    clean[4] = 0.0                                                            # silent reference -> scorer raises -> -1
    with MetricLabelPipeline(_toy_score, workers=3) as pipe:
        h = pipe.submit(clean, est)
        lab = pipe.result(h)
        assert lab.dtype == torch.float32 and lab.shape == (6,)
        want = [(_toy_score(16000, clean[b, :3900].numpy(), est[b].numpy()) - 1) / 3.5 if b != 4 else (-1.0 - 1.0) / 3.5 for b in range(6)]
        assert torch.allclose(lab, torch.tensor(want, dtype=torch.float32), atol=1e-6)
        assert list(pipe.failed(h)) == [False, False, False, False, True, False]
        # submit does not wait for the scorer
        slow = lambda sr, c, n: (time.sleep(0.2), 2.0)[1]
        pipe.score_fn = slow
        t0 = time.perf_counter()
        h2 = pipe.submit(clean, est)
        assert time.perf_counter() - t0 < 0.1
        assert torch.allclose(pipe.result(h2), torch.full((6,), (2.0 - 1) / 3.5))
    # the synchronous reference-shaped call
    out = batch_pesq(list(clean[:3].numpy()), list(clean[:3].numpy()), score_fn=_toy_score, device="cpu")
    assert torch.allclose(out, torch.ones(3), atol=1e-6)


def test_metric_label_pipeline_needs_a_scorer_without_pesq():
    import importlib.util
    from se_b200 import MetricLabelPipeline
    if importlib.util.find_spec("pesq") is None:
        with pytest.raises(ImportError):
            MetricLabelPipeline()


def test_discriminator_state_dict_and_shapes():
    """models/discriminator.py:35-62: same keys as the reference module (spectral-norm parametrisation included), one score in (0, 1) per pair"""
    from se_b200.discriminator import Discriminator
    d = Discriminator(ndf=16)
    keys = set(d.state_dict().keys())
    for i in (0, 3, 6, 9):
        assert {f"layers.{i}.weight_orig", f"layers.{i}.weight_u", f"layers.{i}.weight_v"} <= keys
    assert any(k.endswith(".slope") for k in keys)
    x = torch.rand(2, 1, 201, 81)
    y = d(x, x)
    assert y.shape == (2, 1) and float(y.min()) > 0 and float(y.max()) < 1


def test_metric_label_pipeline_process_backend():
    """worker processes (spawn) like the reference's joblib pool: same labels as the in-process scorer"""
    from se_b200 import MetricLabelPipeline
    from se_b200.metric_labels import log_spectral_score
    g = torch.Generator().manual_seed(1)
    clean = torch.randn(3, 8000, generator=g)
    est = clean + 0.2 * torch.randn(3, 8000, generator=g)
    with MetricLabelPipeline(log_spectral_score, workers=1, backend="process") as pipe:
        pipe.warm_up()
        lab = pipe.result(pipe.submit(clean, est))
    want = torch.tensor([(log_spectral_score(16000, clean[b].numpy(), est[b].numpy()) - 1) / 3.5 for b in range(3)], dtype=torch.float32)
    assert torch.allclose(lab, want, atol=1e-6)
