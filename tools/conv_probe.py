"""Timing probe for the tap-reuse conv kernel: which path (A cp.async / W bulk / MMA) bounds a stage."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200
from se_b200 import ops, packing
from se_b200._lib import EPI_BIAS, LOAD_CONV_SPLIT
B, T, F = 16, 641, 201
for layer in (1, 4):
    slots = [torch.randn(B * T * F, 64, device="cuda") for _ in range(layer)]
    sp = []
    for t in slots:
        y = torch.empty(B * T * F, 2, 64, device="cuda", dtype=torch.bfloat16); ops.split_planes(t, y); sp.append(y)
    w = torch.randn(64, 64 * layer, 2, 3) * 0.05
    pw = packing.pack_weight(packing.conv_weight_matrix(w), 64, torch.zeros(64)).to("cuda")
    out = torch.empty(B * T * F, 64, device="cuda")
    for mode, name in ((0.0, "full"), (1.0, "no A loads"), (2.0, "no W loads"), (3.0, "1/3.. MMAs only first tap")):
        def run():
            ops.gemm(loader=LOAD_CONV_SPLIT, epilogue=EPI_BIAS, M=B * T * F, w=pw, a=sp, out=out, ldo=64, engine="tcgen05", alpha=mode,
                     conv=dict(B=B, T=T, Fin=F, Fout=F, taps_t=2, dil=2 ** (layer - 1), stride_f=1, nslots=layer))
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tiles = (B * T * (F + 1) + 127) // 128
        print(f"layer {layer} {name:28s} {ms:7.3f} ms  -> {ms*1e-3*1.9e9/(tiles/148)/(2*layer):8.0f} cycles per stage per SM")
