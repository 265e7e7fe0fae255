"""Event-timed token GEMMs at the bench size (M = 64 x 641 x 101 tokens) and, with a TG_TRACE build (tools/build_variant_lib.sh tgtrace tok_gemm.cu
-DTG_TRACE=70, SEB200_LIB_SUFFIX=tgtrace), the per-tile clock64 stage trace of one CTA.  usage: tok_trace.py [qkv|glu|out|pw2]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, packing, _lib
from se_b200._lib import EPI_GLU, EPI_QKV_F16, EPI_RESID, LOAD_ROWS, LOAD_ROWS_LN, LOAD_ROWS_F16
torch.manual_seed(0)
dev = "cuda"
M = int(os.environ.get("PROF_B", "64")) * 641 * 101
which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
x = torch.randn(M, 64, device=dev)
g, be = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
if which == "qkv":
    w = packing.pack_weight(torch.randn(192, 64) * 0.17, 192, None).to(dev)
    out = torch.empty(M, 192, device=dev, dtype=torch.float16)
    run = lambda: ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_QKV_F16, M=M, w=w, a=[x], lda=64, ln=(g, be), out=out, ldo=192, engine="tcgen05")
    gb = (4 * 64 + 2 * 192) * M / 1e9
elif which == "glu":
    wi, bi = packing.glu_interleave(torch.randn(256, 64) * 0.17, torch.randn(256) * 0.1)
    w = packing.pack_weight(wi, 256, bi).to(dev)
    out = torch.empty(M, 128, device=dev)
    run = lambda: ops.gemm(loader=LOAD_ROWS_LN, epilogue=EPI_GLU, M=M, w=w, a=[x], lda=64, ln=(g, be), out=out, ldo=128, engine="tcgen05")
    gb = (4 * 64 + 4 * 128) * M / 1e9
elif which == "out":
    w = packing.pack_weight(torch.randn(64, 64) * 0.17, 64, torch.randn(64) * 0.1).to(dev)
    res = torch.randn(M, 64, device=dev)
    run = lambda: ops.gemm(loader=LOAD_ROWS, epilogue=EPI_RESID, M=M, w=w, a=[x], lda=64, resid=res, ldr=64, out=res, ldo=64, engine="tcgen05")
    gb = 3 * 4 * 64 * M / 1e9
elif which == "pw2":
    w = packing.pack_weight(torch.randn(64, 128) * 0.12, 64, torch.randn(64) * 0.1).to(dev)
    xh = torch.randn(M, 128, device=dev).to(torch.float16)
    res = torch.randn(M, 64, device=dev)
    run = lambda: ops.gemm(loader=LOAD_ROWS_F16, epilogue=EPI_RESID, M=M, w=w, a=[xh], lda=128, resid=res, ldr=64, out=res, ldo=64, engine="tcgen05")
    gb = (2 * 128 + 2 * 4 * 64) * M / 1e9
else:
    raise SystemExit("unknown GEMM")
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{which}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic")
lib = _lib.load()
if hasattr(lib, "seb200_tg_trace"):
    buf = (C.c_longlong * (4 * 24 * 4))()
    lib.seb200_tg_trace.argtypes = [C.POINTER(C.c_longlong)]
    lib.seb200_tg_trace(buf)
    tr = torch.tensor(list(buf)).view(4, 24, 4)
    t0 = int(tr[tr > 0].min())
    names = ["row warp 0: want x | got x | stats done + XA free | XA stored", "MMA issuer: want XA | got XA | ACC free | committed",
             "epilogue warp 0: want ACC | got ACC | ACC read | tile done", "copy warp: want slot | got slot | issued"]
    for r in range(4):
        print(names[r])
        for it in range(4, 20):
            print("  ", it, [int(v) - t0 if int(v) else -1 for v in tr[r, it]])
