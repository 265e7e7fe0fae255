"""Per-CTA durations and inter-CTA gaps of the tcgen05 attention kernel (T6_CTATIME build: tools/build_variant_lib.sh ctatime attention_tc.cu -DT6_CTATIME,
SEB200_LIB_SUFFIX=ctatime).  n = 641, 6464 sequences: 51712 CTAs, four query blocks per (sequence, head pair)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, _lib
torch.manual_seed(0)
dev = "cuda"
B, T, Fh = 64, int(os.environ.get("PROF_T", "641")), 101
M = B * T * Fh
inp_h = (torch.randn(M, 192, device=dev) * 0.7).to(torch.float16)
emb = torch.randn(1025, 16, device=dev); emb_h = ops.pack_rel_pos(emb)
seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh)
out = torch.zeros(M, 64, device=dev)
for _ in range(3):
    ops.attention(inp_h, emb, seq, out, 3, emb_h)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((65536, 3), dtype=np.int64)
lib.seb200_t6_cta.argtypes = [C.c_void_p]
lib.seb200_t6_cta(buf.ctypes.data)
nqb = (((T - 1) if (os.environ.get("FRINGE") and T % 64 == 1) else T) + 191) // 192
ncta = min(65536, B * Fh * 2 * nqb)
d = buf[:ncta]
dur = d[:, 1] - d[:, 0]
qb = np.arange(ncta) % nqb
print("CTAs", ncta, "query blocks per (sequence, head pair)", nqb)
for q in range(nqb):
    x = dur[qb == q]
    print(f"  query block {q}: mean duration {x.mean():9.0f} clk  (min {x.min()}, max {x.max()})")
gaps, busy, span = [], [], []
for sm in np.unique(d[:, 2]):
    c = d[d[:, 2] == sm]
    c = c[np.argsort(c[:, 0])]
    gaps.append((c[1:, 0] - c[:-1, 1]))
    busy.append((c[:, 1] - c[:, 0]).sum())
    span.append(c[-1, 1] - c[0, 0])
g = np.concatenate(gaps)
print(f"SMs {len(busy)}; per SM: CTAs {ncta / len(busy):.1f}, busy {np.mean(busy):.3e} clk of span {np.mean(span):.3e} ({np.mean(busy) / np.mean(span) * 100:.1f} %)")
print(f"gap between one CTA's exit and the next CTA's entry on the same SM: mean {g.mean():.0f} clk, median {np.median(g):.0f}, p90 {np.percentile(g, 90):.0f}")
