"""Variant 3 (tcgen05 attention) vs the fp32 kernel (variant 1) and variant 0: error and time per launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops

torch.manual_seed(0)
dev = "cuda"
def run(name, B, T, Fh, axis, variants=(0, 3), check=True, reps=5):
    M = B * T * Fh
    qkv = (torch.randn(M, 192, device=dev) * 1.5)
    emb = torch.randn(1025, 16, device=dev)
    emb_p = ops.pack_rel_pos(emb)
    inp_h = torch.cat([qkv[:, :64] * (0.25 * 1.4426950408889634), qkv[:, 64:]], 1).to(torch.float16).contiguous()
    seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh) if axis == "time" else ops.make_seq(B * T, Fh, 1, Fh, 1)
    ref = None
    if check:
        ref = torch.zeros(M, 64, device=dev)
        ops.attention(qkv, emb, seq, ref, 1)
    for v in variants:
        out = torch.zeros(M, 64, device=dev)
        for _ in range(2):
            ops.attention(inp_h, emb, seq, out, v, emb_p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.attention(inp_h, emb, seq, out, v, emb_p)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        err = float((out - ref).abs().max() / ref.abs().max()) if check else float("nan")
        rms = float((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()) if check else float("nan")
        print(f"{name:28s} variant {v}: {ms:8.3f} ms  max-err/peak {err:.2e}  rel-rms {rms:.2e}", flush=True)
        if check and v == 3 and not (err < 1e-2):
            d = (out - ref).abs().view(-1, 4, 16).amax(-1)      # per (token, head)
            bad = (d > 1e-2 * ref.abs().max()).nonzero()
            print("   bad (token, head) count", bad.shape[0], "first", bad[:12].tolist(), flush=True)

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
if mode == "small":
    run("freq B=1 T=2 Fh=64", 1, 2, 64, "freq")
    run("freq B=1 T=3 Fh=101", 1, 3, 101, "freq")
    run("time B=1 T=130 Fh=3", 1, 130, 3, "time")
    run("time B=1 T=641 Fh=5", 1, 641, 5, "time")
    run("time B=1 T=1300 Fh=2", 1, 1300, 2, "time")
else:
    run("time B=8 T=641 Fh=101", 8, 641, 101, "time")
    run("freq B=8 T=641 Fh=101", 8, 641, 101, "freq")
    run("time B=64 T=641 Fh=101", 64, 641, 101, "time", check=False)
    run("freq B=64 T=641 Fh=101", 64, 641, 101, "freq", check=False)
    run("time B=1 T=4801 Fh=101", 1, 4801, 101, "time", check=False, reps=2)
