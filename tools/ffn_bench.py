"""The fused feed-forward kernel alone at configs[1] size (4.14 M tokens): ms per launch, both forms (plain / post-norm + residual)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200  # noqa: E402
from se_b200 import ops  # noqa: E402
import synth as weights  # noqa: E402

M = 64 * 641 * 101
m = se_b200.TSCNet()
m.load_state_dict(weights.synth_state_dict(0))
m = m.cuda().eval()
P = m.packed()
p = "TSCB_1.time_conformer"
x, y, r = (torch.randn(M, 64, device="cuda") for _ in range(3))


def run(post):
    if post:
        ops.ffn_fused(x, y, P[f"{p}.ff2.ln"], P[f"{p}.ff2.w1"], P[f"{p}.ff2.w2"], 0.5, post=P[f"{p}.post_norm"], resid2=r)
    else:
        ops.ffn_fused(x, y, P[f"{p}.ff1.ln"], P[f"{p}.ff1.w1"], P[f"{p}.ff1.w2"], 0.5)


out = {}
for post in (False, True):
    for _ in range(5):
        run(post)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run(post)
    e1.record()
    torch.cuda.synchronize()
    out["post" if post else "plain"] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out))
