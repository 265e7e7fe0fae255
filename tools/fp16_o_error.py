"""Experiment (CPU, oracle): the attention output o (before the out-projection) stored in fp16, as the CUDA path would write it between the attention
kernel and the out-projection GEMM: whole-path waveform error against the reference goldens."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, synth
from oracle import tscnet_oracle as O
import torch.nn.functional as F
from conftest import load_golden
orig_linear = F.linear


def attention_h(x, sd, p, chunk=0, tr=None):
    S, n, _ = x.shape
    h = O._ln(x, sd, p + ".norm")
    q = F.linear(h, sd[p + ".fn.to_q.weight"]); kv = F.linear(h, sd[p + ".fn.to_kv.weight"])
    k, v = kv[..., :64], kv[..., 64:]
    split = lambda t: t.reshape(S, n, 4, 16).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    dots = torch.matmul(q, k.transpose(-1, -2)) * 0.25
    pos = torch.arange(n)
    dist = (pos[:, None] - pos[None, :]).clamp(-512, 512) + 512
    dots = dots + torch.einsum("bhnd,nrd->bhnr", q, sd[p + ".fn.rel_pos_emb.weight"][dist]) * 0.25
    o = torch.matmul(dots.softmax(-1), v).permute(0, 2, 1, 3).reshape(S, n, 64)
    o = o.half().float()
    return F.linear(o, sd[p + ".fn.to_out.weight"], sd[p + ".fn.to_out.bias"])


for name in ["speech_b2_L8000", "speech_b1_L16000"]:
    g = load_golden(name)
    sd = synth.synth_state_dict(int(g["weight_seed"]))
    noisy = torch.from_numpy(g["noisy"]); ref = torch.from_numpy(g["enhanced"])
    keep = O.attention
    O.attention = attention_h
    with torch.no_grad():
        y = O.predict(noisy, sd)
    O.attention = keep
    print(name, "o16 err %.2e" % float((y - ref).abs().max() / ref.abs().max()))
