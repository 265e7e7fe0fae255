"""Small-shape launches of the default time-axis attention kernel (attention_tc3_kernel, seb200_attention variant 3) for
compute-sanitizer racecheck / memcheck: n = 150 (one partial group), 641 (16-key tail body), 1400 (far-field tiles)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se_b200
from se_b200 import ops

for (nseq, n) in [(2, 150), (1, 641), (1, 1400), (3, 64), (1, 193)]:
    g = torch.Generator().manual_seed(n)
    qkv = (torch.randn(nseq * n, 192, generator=g) * 1.5)
    qkv[:, :64] *= 0.25 * 1.4426950408889634
    emb = torch.randn(1025, 16, generator=g)
    out = torch.zeros(nseq * n, 64, device="cuda")
    seq = ops.make_seq(nseq, n, 1, n, 1)
    ops.attention(qkv.to(torch.float16).cuda(), emb.cuda(), seq, out, 3)
    torch.cuda.synchronize()
    print("n", n, "ok", bool(torch.isfinite(out).all()))
