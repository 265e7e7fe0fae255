"""Event-timed depthwise conv + BN + Swish launches at the bench shapes (fp32 in, fp16 out), both axes; SEB200_DWCONV_CPASYNC=1 selects the cp.async kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops
torch.manual_seed(0)
dev = "cuda"
B, T, Fh = int(os.environ.get("PROF_B", "64")), 641, 101
M = B * T * Fh
x = torch.randn(M, 128, device=dev)
w = torch.randn(31, 128, device=dev) * 0.2
sc, sh = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.1
for axis in ("time", "freq"):
    seq = ops.make_seq(B * Fh, T, Fh, T * Fh, Fh) if axis == "time" else ops.make_seq(B * T, Fh, 1, Fh, 1)
    for dt in (torch.float16, torch.float32):
        y = torch.empty(M, 128, device=dev, dtype=dt)
        for _ in range(3):
            ops.dwconv_bn_swish(x, seq, w, sc, sh, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.dwconv_bn_swish(x, seq, w, sc, sh, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = (4.0 + y.element_size()) * x.numel() / 1e9
        print(f"{axis} out={str(dt)[6:]:8s} {ms:7.3f} ms  {gb / ms * 1e3:7.0f} GB/s", flush=True)
