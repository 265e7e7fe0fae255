"""One network evaluation of the diffusion variant at configs[1] shape (64 x 4 s) for the ncu launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
from se_b200 import tsc_diffusion  # noqa: E402
import synth as weights  # noqa: E402

m = tsc_diffusion.TSCNet(64, 201, noise_schedule=[0.0] * 50)
m.load_state_dict(weights.synth_state_dict(0, spec=weights.tsc_diffusion_spec()))
m = m.cuda().eval()
x, c = weights.synth_wave(64, 64000, seed=1234, kind="speech")
sx, sn = se_b200.compressed_stft(x.cuda()), se_b200.compressed_stft(c.cuda())
torch.cuda.synchronize()
m(sx, sn, torch.tensor([7], device="cuda"))
torch.cuda.synchronize()
