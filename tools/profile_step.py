"""Two passes of the hot path at BASELINE configs[1] (64 x 4 s) for ncu: pass 1 warms up, pass 2 is the one profiled."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
import synth as weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
m = se_b200.TSCNet()
m.load_state_dict(weights.synth_state_dict(0))
m = m.cuda().eval()
enh = se_b200.EnhancerB200(m)
x, _ = weights.synth_wave(B, L, seed=1234, kind="speech")
x = x.cuda()
for _ in range(2):
    enh(x)
    torch.cuda.synchronize()
