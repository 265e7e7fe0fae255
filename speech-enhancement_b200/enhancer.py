"""Wave -> wave enhancement with the semantics of the reference's ``predict()`` (inference_gan.py:75-100), batched.

    enh = EnhancerB200(model)            # model: se_b200.TSCNet (weights loaded, on a CUDA device)
    y = enh(noisy)                       # (B, L) fp32 CUDA -> (B, L)
    y = enh.predict(noisy_numpy_1d)      # the reference's call shape: 1-D numpy in, 1-D numpy out

Per utterance: c = sqrt(L / sum x^2); x*c; wrap-pad to a multiple of 100; compressed STFT; generator;
decompress + iSTFT; /c; trim.  STFT, generator and iSTFT share buffers: the STFT epilogue writes the generator's
3-channel input directly and the generator's output feeds the iDFT rows without a layout change.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import dsp, ops
from .generator import TSCNet


class EnhancerB200(nn.Module):
    def __init__(self, model: TSCNet, n_fft: int = 400, hop: int = 100, use_cuda_graph: bool = False):
        """``use_cuda_graph``: capture the ~135 launches of one forward per (B, L) shape and replay them -- removes the
        host launch overhead that dominates small batches (1 x 2 s: 2.6 ms eager).  The result is returned as a copy."""
        super().__init__()
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self.graph_kernel_nodes = 0
        if n_fft != dsp.N_FFT or hop != dsp.HOP:
            raise ValueError("only the reference configuration N_FFT=400, HOP_SAMPLES=100 is implemented")
        self.model = model
        # DFT / iDFT run on tcgen05 with THREE bf16 planes per operand (six products): |X|^0.3 amplifies operand rounding
        # on near-zero bins, and the two-plane split that serves the network left 6e-4 worst-bin error there (measured).
        self.dft_engine = dsp.DFT_ENGINE

    @torch.no_grad()
    def forward(self, noisy: torch.Tensor, stages=None) -> torch.Tensor:
        if not noisy.is_cuda:
            raise RuntimeError("EnhancerB200 has no CPU path: move the waveform to the GPU first")
        x = noisy.to(torch.float32).contiguous()
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.shape[0] == 0:                               # an empty shard (more ranks than utterances): nothing to launch
            return x.new_zeros(x.shape)
        if x.shape[1] <= dsp.N_FFT // 2:
            raise RuntimeError(f"utterance of {x.shape[1]} samples: torch.stft's reflect padding (n_fft/2 = 200) needs more than 200 "
                               "samples, as in the reference (core/function.py:690)")
        with torch.cuda.device(x.device):                 # launches follow the input's device
            if self.use_cuda_graph and stages is None:
                return self._forward_graphed(x)
            return self._forward_eager(x, stages)

    def _forward_graphed(self, x: torch.Tensor) -> torch.Tensor:
        key = (str(x.device), tuple(x.shape), self.model._version_key(), self.model.engine, self.dft_engine)
        entry = self._graphs.get(key)
        if entry is None:
            for _ in range(2):                       # warm-up: packs weights, sizes workspaces, sets kernel attributes
                self._forward_eager(x, None)
            torch.cuda.synchronize(x.device)
            static_in = x.clone()
            graph = torch.cuda.CUDAGraph()
            from . import _lib
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                static_out = self._forward_eager(static_in, None)
            self.graph_kernel_nodes = _lib.launch_count() - n0      # kernels replayed per call (the C-ABI counter only sees the capture)
            if len(self._graphs) > 8:
                self._graphs.clear()
            # the captured kernels hold raw pointers into the model's workspace tensors and packed weights (allocated by the eager
            # warm-up, outside the graph's private pool): the cache entry keeps both alive for as long as the graph can be replayed,
            # whatever TSCNet.workspace() / packed() evict in the meantime
            B, L = x.shape
            T = int(math.ceil(L / dsp.HOP)) + 1
            keep = (self.model.workspace(B, T, x.device), self.model.packed())
            entry = self._graphs[key] = (graph, static_in, static_out, keep)
        graph, static_in, static_out, _keep = entry
        static_in.copy_(x)
        graph.replay()
        return static_out.clone()

    def _forward_eager(self, x: torch.Tensor, stages=None) -> torch.Tensor:
        B, L = x.shape
        Lp = int(math.ceil(L / dsp.HOP)) * dsp.HOP
        T = Lp // dsp.HOP + 1
        eng = self.dft_engine
        xpad, c = ops.rms_pad(x, Lp, normalize=True)
        in3 = dsp.stft_in3(xpad, T, eng)
        if stages is not None:
            stages["in3"] = in3
        est = self.model.forward_in3(in3, stages)
        z = torch.empty(B * T, dsp.LDZ, device=x.device, dtype=torch.float32)
        ops.decompress_rows(est, z)
        y = dsp.istft_rows(z, B, T, c, eng)
        return y[:, :L]

    def predict(self, noisy_signal: np.ndarray) -> np.ndarray:
        """inference_gan.predict(model, config, noisy_signal): 1-D numpy -> 1-D numpy of the same length."""
        dev = next(self.model.parameters()).device
        x = torch.from_numpy(np.ascontiguousarray(noisy_signal, dtype=np.float32)).to(dev)
        return self.forward(x.unsqueeze(0))[0].cpu().numpy()
