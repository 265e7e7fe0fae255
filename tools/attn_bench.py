"""Event-timed attention launches at the bench shapes: time axis (n = 641, 6464 sequences), frequency axis (n = 101, 41024 sequences),
30 s clips (n = 4801, B = 4).  usage: attn_bench.py [variant ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, _lib
if os.environ.get("SEB200_LIB_SUFFIX"):
    _lib._LIB_PATH = os.path.join(os.path.dirname(_lib._LIB_PATH), "libseb200_" + os.environ["SEB200_LIB_SUFFIX"] + ".so")
torch.manual_seed(0)
dev = "cuda"
variants = [int(v) for v in sys.argv[1:]] or [0, 3]
emb = torch.randn(1025, 16, device=dev); emb_h = ops.pack_rel_pos(emb)


def run(name, B, T, Fh, axis, v, reps=10):
    M = B * T * Fh
    inp_h = (torch.randn(M, 192, device=dev) * 0.7).to(torch.float16)
    seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh) if axis == "time" else ops.make_seq(B * T, Fh, 1, Fh, 1)
    out = torch.zeros(M, 64, device=dev)
    for _ in range(3):
        ops.attention(inp_h, emb, seq, out, v, emb_h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.attention(inp_h, emb, seq, out, v, emb_h)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 96.0 * 4 * seq.nseq * T * T if axis == "time" else 96.0 * 4 * seq.nseq * Fh * Fh
    print(f"{name:28s} variant {v}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    return out


for v in variants:
    run("time n=641 B=64", 64, 641, 101, "time", v)
    if os.environ.get("ATTN_BENCH_SHORT"):
        continue
    run("freq n=101 B=64", 64, 641, 101, "freq", v)
    run("time n=4801 B=4", 4, 4801, 101, "time", v, reps=3)
    run("time n=1001 B=16", 16, 1001, 101, "time", v, reps=3)
if len(variants) > 1:
    a = run("check", 2, 641, 101, "time", variants[0], 1)
    b = run("check", 2, 641, 101, "time", variants[-1], 1)
    print("max |a - b| / max |a|:", float((a - b).abs().max() / a.abs().max()))
