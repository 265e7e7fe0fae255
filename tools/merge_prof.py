"""The two MergeBlock launches at configs[1] size (4.14 M tokens) for ncu: three rounds, profile the last."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
from se_b200 import ops, tsc_diffusion  # noqa: E402
from se_b200._lib import EPI_GATE, EPI_RESID_SCALE, LOAD_ROWS, LOAD_ROWS2  # noqa: E402
import synth as weights  # noqa: E402

M = 64 * 641 * 101
m = tsc_diffusion.TSCNet(64, 201, noise_schedule=[0.0] * 50)
m.load_state_dict(weights.synth_state_dict(0, spec=weights.tsc_diffusion_spec()))
m = m.cuda().eval()
P = m.packed()
x, cond, g = (torch.randn(M, 64, device="cuda") for _ in range(3))
rowbias = torch.randn(64, 128, device="cuda")
for _ in range(3):
    ops.gemm(loader=LOAD_ROWS2, epilogue=EPI_GATE, M=M, w=P["merge_block.gate"], a=[x, cond], lda=64, out=g, ldo=64, resid=rowbias, ldr=641 * 101)
    ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID_SCALE, M=M, w=P["merge_block.out"], a=[g], lda=64, out=x, ldo=64, resid=x, ldr=64, alpha=1 / math.sqrt(2))
    torch.cuda.synchronize()
