"""Per-tile clock64 trace of one CTA of the tcgen05 attention kernel (debug build libseb200_trace.so, tools/build_trace_lib.sh)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, _lib
_lib._LIB_PATH = os.path.join(os.path.dirname(_lib._LIB_PATH), "libseb200_" + os.environ.get("SEB200_LIB_SUFFIX", "trace") + ".so")
torch.manual_seed(0)
dev = "cuda"
B, T, Fh = int(os.environ.get("PROF_B", "64")), int(os.environ.get("PROF_T", "641")), 101
M = B * T * Fh
inp_h = (torch.randn(M, 192, device=dev) * 0.7).to(torch.float16)
emb = torch.randn(1025, 16, device=dev); emb_h = ops.pack_rel_pos(emb)
seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh)
out = torch.zeros(M, 64, device=dev)
for _ in range(3):
    ops.attention(inp_h, emb, seq, out, int(os.environ.get("TRACE_VARIANT", "3")), emb_h)
torch.cuda.synchronize()
buf = (C.c_longlong * (6 * 16 * 8))()
lib = _lib.load()
lib.seb200_t6_trace.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.seb200_t6_trace(buf))
tr = torch.tensor(list(buf)).view(6, 16, 8)
t0 = int(tr[tr > 0].min())
nt = (T + 63) // 64
for g in range(3):
    print(f"softmax group {g} warp 0 (clk since start): tile: wantS gotS R_read S_read max gotO exp_done P_arrive")
    for t in range(nt):
        print(" ", t, [int(v) - t0 if int(v) else -1 for v in tr[g, t]])
    d = [[int(tr[g, t, e + 1]) - int(tr[g, t, e]) for e in range(7)] + [int(tr[g, t + 1, 0]) - int(tr[g, t, 0])] for t in range(2, nt - 2)]
    print("  mean phase lengths over the inner tiles (7 phases, tile period):", [sum(c) // len(c) for c in zip(*d)])
print("MMA 1 issuer: tile: [ready(g) committed(g)] x 3, full_bar")
for t in range(nt):
    print(" ", t, [int(v) - t0 if int(v) else -1 for v in tr[3, t, :7]])
print("MMA 3 issuer: tile: [gotP(g) committed(g)] x 3")
for t in range(nt):
    print(" ", t, [int(v) - t0 if int(v) else -1 for v in tr[4, t, :6]])
