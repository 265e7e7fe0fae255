"""Same-box A/B of TSCNet.overlap_decoders (complex decoder on a side stream) at configs[1]: step time off / on, bitwise equality."""
import json
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


def timed(flag, reps=5):
    m.overlap_decoders = flag
    for _ in range(3):
        y = enh(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = enh(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, y.clone()


res = {}
for rnd in range(2):
    for flag in (False, True):
        ms, y = timed(flag)
        res.setdefault(str(flag), []).append(round(ms, 2))
        if flag:
            res["equal"] = bool(torch.equal(y, y_off))
        else:
            y_off = y
print(json.dumps(res))
