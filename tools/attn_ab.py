"""A/B of the attention kernels at BASELINE configs[1] shapes: error vs the fp32 kernel (variant 1) and time per launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops

torch.manual_seed(0)
dev = "cuda"
def run(name, B, T, Fh, axis, variants=(2, 0), check=True):
    M = B * T * Fh
    qkv = (torch.randn(M, 192, device=dev) * 1.5)
    emb = torch.randn(1025, 16, device=dev)
    emb_h = emb.to(torch.float16).contiguous(); emb_p = ops.pack_rel_pos(emb)
    inp_h = torch.cat([qkv[:, :64] * (0.25 * 1.4426950408889634), qkv[:, 64:]], 1).to(torch.float16).contiguous()
    seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh) if axis == "time" else ops.make_seq(B * T, Fh, 1, Fh, 1)
    ref = None
    if check:
        ref = torch.zeros(M, 64, device=dev)
        ops.attention(qkv, emb, seq, ref, 1)
    for v in variants:
        out = torch.zeros(M, 64, device=dev)
        for _ in range(2):
            ops.attention(inp_h, emb, seq, out, v, emb_p if v == 0 else emb_h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.attention(inp_h, emb, seq, out, v, emb_p if v == 0 else emb_h)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        err = float((out - ref).abs().max() / ref.abs().max()) if check else float("nan")
        rms = float((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()) if check else float("nan")
        print(f"{name:28s} variant {v}: {ms:8.3f} ms  max-err/peak {err:.2e}  rel-rms {rms:.2e}", flush=True)

run("time B=8 T=641 Fh=101", 8, 641, 101, "time")
run("freq B=8 T=641 Fh=101", 8, 641, 101, "freq")
run("time B=64 T=641 Fh=101", 64, 641, 101, "time", check=False)
run("freq B=64 T=641 Fh=101", 64, 641, 101, "freq", check=False)
run("time B=1 T=4801 Fh=101", 1, 4801, 101, "time", check=False)
