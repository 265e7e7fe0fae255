"""One reverse-diffusion network evaluation at BASELINE configs[1] shape: tsc_diffusion.TSCNet.forward on
64 x 4 s spectrogram pairs; prints the device time per evaluation and the per-kernel split (CUDA events on the launching stream)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
from se_b200 import ops, tsc_diffusion  # noqa: E402
import synth as weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
m = tsc_diffusion.TSCNet(64, 201, noise_schedule=[0.0] * 50)
m.load_state_dict(weights.synth_state_dict(0, spec=weights.tsc_diffusion_spec()))
m = m.cuda().eval()
x, c = weights.synth_wave(B, L, seed=1234, kind="speech")
sx, sn = se_b200.compressed_stft(x.cuda()), se_b200.compressed_stft(c.cuda())
step = torch.tensor([7], device="cuda")
for _ in range(3):
    m(sx, sn, step)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m(sx, sn, step)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
with ops.profile() as prof:
    m(sx, sn, step)
summ = prof.summary()
M = B * (L // 100 + 1) * 101
out = {"ms_per_eval": ms, "audio_s_per_s_per_eval": B * L / 16000 / (ms * 1e-3),
       "kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])},
       "merge_gate_GBps": 3 * M * 256 / (summ["merge_gate"]["ms"] / summ["merge_gate"]["launches"] * 1e-3) / 1e9,
       "merge_out_GBps": 3 * M * 256 / (summ["merge_out"]["ms"] / summ["merge_out"]["launches"] * 1e-3) / 1e9}
print(json.dumps(out))
