"""Weight packing for the GEMM engine (host side, runs on CPU tensors).

Every dense contraction on the path is ``C[M,N] = A[M,K] . W[N,K]^T``.  For one
logical weight matrix W (N x K, fp32) two images are produced:

* ``w_tc``  (tcgen05 main loop): K padded to a multiple of 64, N padded to
  ``ntiles * ntile``.  W is split into bf16 ``hi = bf16(W)`` and
  ``lo = bf16(W - hi)``.  For every (n-tile j, k-chunk kc) the image holds the
  ``hi`` block then the ``lo`` block, each ``ntile`` rows x 128 bytes in the
  UMMA K-major SWIZZLE_128B canonical layout: the 16-byte chunk ``c`` of row
  ``r`` sits at byte ``r*128 + ((c ^ (r & 7)) << 4)``.  One ``cp.async.bulk``
  of ``2 * ntile * 128`` bytes brings a stage's weights into shared memory
  ready for ``tcgen05.mma``.
* ``w_simt`` (fp32 FFMA main loop): ``[Kpad][npad]`` fp32 (K-major rows,
  npad a multiple of 64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

BK = 64


@dataclass
class PackedWeight:
    N: int
    K: int                 # padded K (multiple of 64)
    tc_ntile: int
    tc_ntiles: int
    simt_npad: int
    w_tc: torch.Tensor     # uint8
    w_simt: torch.Tensor   # float32 [K][simt_npad]
    bias: Optional[torch.Tensor]
    planes: int = 2        # bf16 planes per operand in w_tc: 2 (hi|lo) or 3 (hi|mid|lo)

    def to(self, device):
        return PackedWeight(self.N, self.K, self.tc_ntile, self.tc_ntiles, self.simt_npad,
                            self.w_tc.to(device), self.w_simt.to(device),
                            None if self.bias is None else self.bias.to(device), self.planes)


def split_bf16(w: torch.Tensor):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def swizzle128(block: torch.Tensor) -> torch.Tensor:
    """block: [rows, 64] bf16 (one 128-byte row each) -> same shape with 16-byte chunks XOR-permuted."""
    rows = block.shape[0]
    b = block.reshape(rows, 8, 8)
    r = torch.arange(rows).view(rows, 1)
    p = torch.arange(8).view(1, 8)
    src = p ^ (r & 7)                                    # position p holds chunk p ^ (r & 7)
    return torch.gather(b, 1, src.view(rows, 8, 1).expand(rows, 8, 8)).reshape(rows, 64)


def split_bf16_3(w: torch.Tensor):
    hi = w.to(torch.bfloat16)
    r = w - hi.to(torch.float32)
    mid = r.to(torch.bfloat16)
    lo = (r - mid.to(torch.float32)).to(torch.bfloat16)
    return hi, mid, lo


def pack_weight(w: torch.Tensor, tc_ntile: int, bias: Optional[torch.Tensor] = None, planes: int = 2) -> PackedWeight:
    """w: [N, K] fp32 (CPU).  planes = 3 packs hi|mid|lo (six-product mode of the engine, used by the DFT bases).
    The images come from the library's own host-side packer (``seb200_pack_weights``, csrc/pack.cu) -- the same entry point a
    non-Python caller of the C ABI uses; ``pack_weight_torch`` is the independent restatement the tests compare it with."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    w = w.detach().to(torch.float32).cpu().contiguous()
    N, K = w.shape
    tcb, sf, kp, ntiles, npad = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.seb200_packed_weight_sizes(N, K, tc_ntile, planes, C.byref(tcb), C.byref(sf), C.byref(kp), C.byref(ntiles), C.byref(npad)),
               "seb200_packed_weight_sizes")
    w_tc = torch.empty(tcb.value, dtype=torch.uint8)
    w_simt = torch.empty(kp.value, npad.value, dtype=torch.float32)
    _lib.check(lib.seb200_pack_weights(w.data_ptr(), N, K, tc_ntile, planes, w_tc.data_ptr(), w_simt.data_ptr()), "seb200_pack_weights")
    b = None if bias is None else bias.detach().to(torch.float32).cpu().contiguous()
    return PackedWeight(N, kp.value, tc_ntile, ntiles.value, npad.value, w_tc, w_simt, b, planes)


def pack_weight_torch(w: torch.Tensor, tc_ntile: int, bias: Optional[torch.Tensor] = None, planes: int = 2) -> PackedWeight:
    """the same images built with torch ops (layout restated independently of csrc/pack.cu; used by tests/test_host.py)"""
    w = w.detach().to(torch.float32).cpu().contiguous()
    N, K = w.shape
    kp = int(math.ceil(K / BK)) * BK
    ntiles = int(math.ceil(N / tc_ntile))
    npad_tc = ntiles * tc_ntile
    wp = torch.zeros(npad_tc, kp, dtype=torch.float32)
    wp[:N, :K] = w
    parts = split_bf16(wp) if planes == 2 else split_bf16_3(wp)
    nkc = kp // BK
    img = torch.empty(ntiles, nkc, planes, tc_ntile, BK, dtype=torch.bfloat16)
    for j in range(ntiles):
        for kc in range(nkc):
            sl = (slice(j * tc_ntile, (j + 1) * tc_ntile), slice(kc * BK, (kc + 1) * BK))
            for pi, part in enumerate(parts):
                img[j, kc, pi] = swizzle128(part[sl])
    w_tc = img.contiguous().view(torch.uint8).reshape(-1)
    npad = int(math.ceil(N / 64)) * 64
    w_simt = torch.zeros(kp, npad, dtype=torch.float32)
    w_simt[:K, :N] = w.t()
    b = None if bias is None else bias.detach().to(torch.float32).cpu().contiguous()
    return PackedWeight(N, kp, tc_ntile, ntiles, npad, w_tc, w_simt.contiguous(), b, planes)


def conv_weight_matrix(w: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [Cout, Cin, kt, kf] -> [Cout, K] with K order (kt, kf, cin): the implicit-GEMM
    loader walks taps outermost and the (newest-first) 64-channel slots of the dense block innermost."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def glu_interleave(w: torch.Tensor, b: torch.Tensor):
    """pointwise Conv1d(64 -> 256) feeding GLU (conformer.py:165-166): rows [value(128) | gate(128)]
    -> rows (value_0, gate_0, value_1, gate_1, ...) so the epilogue sees each pair in one thread."""
    half = w.shape[0] // 2
    wi = torch.stack([w[:half], w[half:]], dim=1).reshape(w.shape[0], -1)
    bi = torch.stack([b[:half], b[half:]], dim=1).reshape(-1)
    return wi, bi


def hamming_periodic(n: int = 400) -> torch.Tensor:
    k = torch.arange(n, dtype=torch.float64)
    return 0.54 - 0.46 * torch.cos(2.0 * math.pi * k / n)


def dft_basis(n_fft: int = 400) -> torch.Tensor:
    """[2*(n_fft/2+1), n_fft]: rows (2k, 2k+1) = window * (cos, -sin)(2 pi k n / n_fft)  (torch.stft, core/function.py:690)."""
    nb = n_fft // 2 + 1
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(nb, dtype=torch.float64).view(-1, 1)
    ang = 2.0 * math.pi * ((k * n) % n_fft) / n_fft
    w = hamming_periodic(n_fft)
    basis = torch.stack([torch.cos(ang) * w, -torch.sin(ang) * w], dim=1).reshape(2 * nb, n_fft)
    return basis.to(torch.float32)


def idft_basis(n_fft: int = 400) -> torch.Tensor:
    """[n_fft, 2*(n_fft/2+1)]: Hermitian-weighted inverse real DFT with the synthesis window folded in
    (torch.istft, core/function.py:701-702).  Columns (2k, 2k+1) multiply (Re Z_k, Im Z_k)."""
    nb = n_fft // 2 + 1
    n = torch.arange(n_fft, dtype=torch.float64).view(-1, 1)
    k = torch.arange(nb, dtype=torch.float64).view(1, -1)
    ang = 2.0 * math.pi * ((n * k) % n_fft) / n_fft
    ck = torch.full((1, nb), 2.0, dtype=torch.float64)
    ck[0, 0] = 1.0
    ck[0, -1] = 1.0
    w = hamming_periodic(n_fft).view(-1, 1)
    re = w * ck * torch.cos(ang) / n_fft
    im = -w * ck * torch.sin(ang) / n_fft
    return torch.stack([re, im], dim=2).reshape(n_fft, 2 * nb).to(torch.float32)


def inv_envelope(T: int, n_fft: int = 400, hop: int = 100) -> torch.Tensor:
    """1 / sum_t w^2[m + n_fft/2 - hop*t] for the hop*(T-1) samples torch.istft keeps (center=True)."""
    w2 = hamming_periodic(n_fft) ** 2
    full = n_fft + hop * (T - 1)
    env = torch.zeros(full, dtype=torch.float64)
    idx = (torch.arange(T) * hop)[:, None] + torch.arange(n_fft)[None, :]
    env.index_add_(0, idx.reshape(-1), w2.to(torch.float64).repeat(T))        # one scatter instead of a Python loop over T frames
    half = n_fft // 2
    return (1.0 / env[half:full - half]).to(torch.float32)
