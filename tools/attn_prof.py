import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops
torch.manual_seed(0)
dev = "cuda"
B, T, Fh = int(os.environ.get('PROF_B', '64')), 641, 101
M = B * T * Fh
v = int(sys.argv[1]) if len(sys.argv) > 1 else 0
axis = sys.argv[2] if len(sys.argv) > 2 else "time"
inp_h = (torch.randn(M, 192, device=dev) * 0.7).to(torch.float16)
emb = torch.randn(1025, 16, device=dev); emb_h = ops.pack_rel_pos(emb) if v in (0, 3) else emb.to(torch.float16).contiguous()
seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh) if axis == "time" else ops.make_seq(B * T, Fh, 1, Fh, 1)
out = torch.zeros(M, 64, device=dev)
for _ in range(3):
    ops.attention(inp_h, emb, seq, out, v, emb_h)
torch.cuda.synchronize()
