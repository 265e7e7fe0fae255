"""Drop-in ``TSCNet`` of the diffusion variant, running on libseb200 (sm_100a).

Mirrors /root/reference/models/tsc_diffusion.py:43-90:

    TSCNet(num_channel=64, num_features=201, noise_schedule=...).forward(x, noisy_spec, diffusion_step)
        -> (final_real, final_imag)

* same constructor arguments and sub-module tree as the reference, hence the same ``state_dict`` keys / shapes:
  the GAN generator's 359 entries plus ``dense_encoder_noisy.*`` and ``merge_block.{diffusion_embedding.projection1,
  diffusion_embedding.projection2, diffusion_projection, merge_diffusion, conditioner_projection, output_residual}.*``;
  the sinusoidal step table is a non-persistent buffer exactly as in models/DiffuSE.py:42.
* ``forward`` takes the two complex64 ``(B, 201, T)`` compressed spectrograms (the current estimate and the conditioning
  noisy utterance) and the diffusion step -- an int or float tensor with one entry, or one per utterance, as
  inference_diffuse.py:251-252 passes it -- and returns two fp32 ``(B, 1, T, 201)`` tensors.

Every stage other than the MergeBlock reuses the generator's kernels (both encoders, the four TSCBs, both decoders).
The MergeBlock (tsc_diffusion.py:16-41), called before each TSCB with the same parameters, is two launches of the persistent
tcgen05 token GEMM plus one tiny launch per forward:

    seb200_diffusion_embed                         step -> d [n, 64] and rowbias = W_m d  (128 -> 512 -> 512 -> 64 MLP)
    SEB_LOAD_ROWS2 + SEB_EPI_GATE      (K = 128)   g = sigmoid(.) * tanh(.) of [W_m | W_c] [x | cond] + b_m + b_c + rowbias
    SEB_LOAD_ROWS  + SEB_EPI_RESID_SCALE (K = 64)  x <- (x + W_o g + b_o) / sqrt(2)

so the 128-wide pre-activation never reaches HBM.  No CPU path and no eager fallback; non-CUDA input raises.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import EPI_GATE, EPI_RESID_SCALE, LOAD_ROWS, LOAD_ROWS2
from .generator import TSCNet as _GanTSCNet
from .generator import _Bag, _dense_encoder
from .packing import glu_interleave, pack_weight


def _step_table(max_steps: int) -> torch.Tensor:
    """DiffusionEmbedding._build_embedding (models/DiffuSE.py:64-69): [max_steps, 128] = (sin | cos)(step * 10^(4 j / 63))."""
    steps = torch.arange(max_steps).unsqueeze(1)
    dims = torch.arange(64).unsqueeze(0)
    table = steps * 10.0 ** (dims * 4.0 / 63.0)
    return torch.cat([torch.sin(table), torch.cos(table)], dim=1)


class TSCNet(_GanTSCNet):
    def __init__(self, num_channel: int = 64, num_features: int = 201, noise_schedule=None):
        if noise_schedule is None:
            raise TypeError("tsc_diffusion.TSCNet needs noise_schedule (its length sizes the step table, tsc_diffusion.py:19)")
        super().__init__(num_channel, num_features)
        ch = num_channel
        self.dense_encoder_noisy = _dense_encoder(ch)
        emb = _Bag()
        emb.register_buffer("embedding", _step_table(len(noise_schedule)), persistent=False)
        emb.projection1 = nn.Linear(128, 512)
        emb.projection2 = nn.Linear(512, 512)
        mb = _Bag()
        mb.diffusion_embedding = emb
        mb.diffusion_projection = nn.Linear(512, ch)
        mb.merge_diffusion = nn.Conv2d(ch, ch * 2, 1)
        mb.conditioner_projection = nn.Conv2d(ch, ch * 2, 1)
        mb.output_residual = nn.Conv2d(ch, ch, 1)
        self.merge_block = mb
        # registration order of the reference (tsc_diffusion.py:46-57), so state_dict() lists the keys in the same order
        order = ["dense_encoder", "dense_encoder_noisy", "merge_block", "TSCB_1", "TSCB_2", "TSCB_3", "TSCB_4", "mask_decoder", "complex_decoder"]
        for k in order:
            self._modules[k] = self._modules.pop(k)
        self.max_steps = len(noise_schedule)
        self._merge = None            # (cond, rowbias, rows_per_group) of the forward in flight

    # ---- packing -------------------------------------------------------------------------------
    def _encoder_names(self):
        return ("dense_encoder", "dense_encoder_noisy")

    def _pack_extra(self, sd, P, dev, device):
        m = "merge_block"
        wm = sd[f"{m}.merge_diffusion.weight"].reshape(128, 64)
        wc = sd[f"{m}.conditioner_projection.weight"].reshape(128, 64)
        # rows (gate_j, filter_j) adjacent: torch.chunk(y, 2, dim=1) takes gate = channels 0..63, filter = 64..127 (tsc_diffusion.py:36)
        wcat, bcat = glu_interleave(torch.cat([wm, wc], dim=1), sd[f"{m}.merge_diffusion.bias"] + sd[f"{m}.conditioner_projection.bias"])
        P[f"{m}.gate"] = pack_weight(wcat, 128, bcat).to(device)
        P[f"{m}.wm_rows"] = dev(wcat[:, :64])
        P[f"{m}.out"] = pack_weight(sd[f"{m}.output_residual.weight"].reshape(64, 64), 64, sd[f"{m}.output_residual.bias"]).to(device)
        for k in ("diffusion_embedding.projection1", "diffusion_embedding.projection2", "diffusion_projection"):
            P[f"{m}.{k}"] = (dev(sd[f"{m}.{k}.weight"]), dev(sd[f"{m}.{k}.bias"]))
        P[f"{m}.table"] = dev(self.merge_block.diffusion_embedding.embedding.detach().to("cpu", torch.float32))

    def workspace(self, B, T, device):
        ws = super().workspace(B, T, device)
        if "cond" not in ws:
            Ph = B * T * ((self.num_features - 1) // 2 + 1)
            ws["cond"] = torch.empty(Ph, 64, device=device, dtype=torch.float32)
        return ws

    # ---- MergeBlock.forward (tsc_diffusion.py:27-41) -------------------------------------------------
    def _before_tscb(self, P, ws, x, i):
        cond, rowbias, rows_per_group = self._merge
        M = x.shape[0]
        g = ws["o"]                                   # [M, 64] gate output (free between TSCBs)
        ops.gemm(loader=LOAD_ROWS2, epilogue=EPI_GATE, M=M, w=P["merge_block.gate"], a=[x, cond], lda=64, out=g, ldo=64,
                 resid=rowbias, ldr=rows_per_group, engine=self.engine, label="merge_gate")
        ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID_SCALE, M=M, w=P["merge_block.out"], a=[g], lda=64, out=x, ldo=64,
                 resid=x, ldr=64, alpha=1.0 / math.sqrt(2.0), engine=self.engine, label="merge_out")

    def _step_bias(self, P, diffusion_step, B: int, device):
        if diffusion_step is None:
            raise RuntimeError("tsc_diffusion.TSCNet.forward needs diffusion_step (tsc_diffusion.py:27)")
        st = torch.as_tensor(diffusion_step, device=device).reshape(-1)
        n = st.numel()
        if n not in (1, B):
            raise RuntimeError(f"diffusion_step must hold 1 or B={B} entries, got {n}")
        st = st.to(torch.float32).contiguous()
        d = torch.empty(n, 64, device=device, dtype=torch.float32)
        rowbias = torch.empty(n, 128, device=device, dtype=torch.float32)
        m = "merge_block"
        ops.diffusion_embed(st, P[f"{m}.table"], *P[f"{m}.diffusion_embedding.projection1"], *P[f"{m}.diffusion_embedding.projection2"],
                            *P[f"{m}.diffusion_projection"], P[f"{m}.wm_rows"], d, rowbias)
        return d, rowbias

    def forward_in3(self, in3: torch.Tensor, noisy_in3: torch.Tensor = None, diffusion_step=None, stages: Optional[dict] = None) -> torch.Tensor:
        """in3 / noisy_in3: [B, T, F, 3] = (|Y|, Re Y, Im Y) of the two compressed spectrograms -> est [B*T, F, 2] (workspace tensor)."""
        if self.training:
            raise RuntimeError("se_b200.tsc_diffusion.TSCNet is in train() mode: only the inference forward is implemented; call .eval()")
        if noisy_in3 is None or noisy_in3.shape != in3.shape:
            raise RuntimeError("tsc_diffusion.TSCNet needs the conditioning spectrogram with the shape of x")
        B, T, F, _ = in3.shape
        if F != self.num_features or F % 2 == 0:
            raise RuntimeError(f"expected {self.num_features} frequency bins, got {F}")
        Fh = (F - 1) // 2 + 1
        dev = in3.device
        P = self.packed()
        ws = self.workspace(B, T, dev)
        d, rowbias = self._step_bias(P, diffusion_step, B, dev)
        cond = self._encode(P, "dense_encoder_noisy", ws, noisy_in3, ws["cond"])      # tsc_diffusion.py:75
        x = self._encode(P, "dense_encoder", ws, in3, ws["x"])                          # :74
        if stages is not None:
            stages["encoder"] = x.view(B, T, Fh, 64).clone()
            stages["encoder_noisy"] = cond.view(B, T, Fh, 64).clone()
            stages["step_projection"] = d.clone()
        self._merge = (cond, rowbias, T * Fh if rowbias.shape[0] == B and B > 1 else B * T * Fh)
        try:
            self._tscbs(P, ws, x, B, T, Fh, stages)                                     # :77-80, merge_block before every TSCB
        finally:
            self._merge = None
        return self._decode(P, ws, x, in3, stages)                                      # :82-90 (mask / phase from x, not from noisy_spec)

    def forward(self, x: torch.Tensor, noisy_spec: torch.Tensor = None, diffusion_step=None):
        """x, noisy_spec: complex64 (B, num_features, T) -> (final_real, final_imag), each fp32 (B, 1, T, F)."""
        if noisy_spec is None:
            raise TypeError("forward(x, noisy_spec, diffusion_step): noisy_spec is required (tsc_diffusion.py:60)")
        if not (x.is_cuda and noisy_spec.is_cuda):
            raise RuntimeError("se_b200.tsc_diffusion.TSCNet has no CPU path: inputs must be CUDA tensors on an sm_100a device")
        if not (x.is_complex() and noisy_spec.is_complex()):
            raise RuntimeError("TSCNet.forward expects complex compressed spectrograms (B, F, T)")
        if self.training:
            raise RuntimeError("se_b200.tsc_diffusion.TSCNet is in train() mode: only the inference forward is implemented; call .eval()")
        with torch.no_grad(), torch.cuda.device(x.device):
            B, F, T = x.shape
            in3 = ops.spec_to_in3(x.to(torch.complex64))
            nin3 = ops.spec_to_in3(noisy_spec.to(torch.complex64))
            est = self.forward_in3(in3, nin3, diffusion_step)
            fr = torch.empty(B, 1, T, F, device=x.device, dtype=torch.float32)
            fi = torch.empty(B, 1, T, F, device=x.device, dtype=torch.float32)
            ops.split_ri(est, fr, fi)
        return fr, fi
