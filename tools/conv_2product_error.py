"""VERDICT r1 item 4 experiment (CPU, oracle): the dilated convs with TWO tensor-core products per logical product -- activations split
bf16 hi + lo as today, weights in a SINGLE 16-bit plane (fp16: 11-bit mantissa, or bf16) -- i.e. a_hi.w + a_lo.w, dropping a_hi.w_lo.
Whole-path waveform error against the reference goldens, next to today's three-product split."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import synth
from oracle import tscnet_oracle as O
from conftest import load_golden

def split2(x):
    hi = x.to(torch.bfloat16).float()
    return hi + (x - hi).to(torch.bfloat16).float()

orig = F.conv2d
def make(wmode):
    def conv2d(x, w, b=None, *a, **k):
        if w.shape[-1] == 3 and w.shape[1] >= 64:       # the (2,3) / (1,3) convs of the dense blocks, conv_2, sub-pixel
            x = split2(x)
            w = {"fp16": w.half().float(), "bf16": w.to(torch.bfloat16).float(), "split": split2(w)}[wmode]
        return orig(x, w, b, *a, **k)
    return conv2d

for name in ["speech_b2_L8000", "speech_b1_L16000", "noise_b1_L4050_wrap"]:
    g = load_golden(name)
    sd = synth.synth_state_dict(int(g["weight_seed"]))
    noisy, ref = torch.from_numpy(g["noisy"]), torch.from_numpy(g["enhanced"])
    for wmode in ("split", "fp16", "bf16"):
        F.conv2d = make(wmode)
        with torch.no_grad():
            y = O.predict(noisy, sd)
        F.conv2d = orig
        print(f"{name:22s} conv weights {wmode:5s}: waveform max-abs / peak {float((y - ref).abs().max() / ref.abs().max()):.2e}", flush=True)
