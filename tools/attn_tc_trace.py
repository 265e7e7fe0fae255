"""Per-tile clock64 trace of one CTA of the tcgen05 attention kernel (debug build libseb200_trace.so, -DT5_TRACE=<block>)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, _lib
_lib._LIB_PATH = os.path.join(os.path.dirname(_lib._LIB_PATH), "libseb200_trace.so")
torch.manual_seed(0)
dev = "cuda"
B, T, Fh = 16, 641, 101
M = B * T * Fh
inp_h = (torch.randn(M, 192, device=dev) * 0.7).to(torch.float16)
emb = torch.randn(1025, 16, device=dev); emb_h = ops.pack_rel_pos(emb)
seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh)
out = torch.zeros(M, 64, device=dev)
for _ in range(3):
    ops.attention(inp_h, emb, seq, out, 3, emb_h)
torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 16 * 8))()
lib = _lib.load()
lib.seb200_t5_trace.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.seb200_t5_trace(buf))
tr = torch.tensor(list(buf)).view(2, 16, 8)
t0 = int(tr[0, 0, 0])
print("softmax warp 0 (rel clk): tile: wantS gotS  F_arrive  preO  gotO  preStWait  P_arrive | control: wantF gotF mma1_issued gotP mma3_issued")
for t in range(11):
    a = [int(v) - t0 if int(v) else -1 for v in tr[0, t, :7]]
    c = [int(v) - t0 if int(v) else -1 for v in tr[1, t, :5]]
    print(t, a, "|", c)
