"""Per-kernel-label time of one generator training step (train-mode forward + backward) at BASELINE configs[4]'s per-GPU shape:
CUDA events around every launch (ops.profile).  python tools/train_prof.py [engine] [B] [seconds]"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se_b200, synth
from se_b200 import ops

engine = sys.argv[1] if len(sys.argv) > 1 else "tcgen05_f32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sec = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
L = int(sec * 16000)
m = se_b200.TSCNet(64, 201)
m.load_state_dict(synth.synth_state_dict(0))
m = m.cuda().train()
se_b200.training._state(m).engine = engine
noisy, _ = synth.synth_wave(B, L, 1234, "speech")
spec = se_b200.compressed_stft(noisy.cuda())
for _ in range(2):
    fr, fi = m(spec)
    (fr.square().mean() + fi.square().mean()).backward()
    m.zero_grad()
torch.cuda.synchronize()
with ops.profile() as prof:
    fr, fi = m(spec)
    torch.cuda.synchronize()
    fwd = prof.summary()
with ops.profile() as prof:
    (fr.square().mean() + fi.square().mean()).backward()
    torch.cuda.synchronize()
    bwd = prof.summary()
out = {}
for name, tab in (("forward", fwd), ("backward", bwd)):
    tot = sum(v["ms"] for v in tab.values())
    print(f"== {name}: {tot:.2f} ms in {sum(v['launches'] for v in tab.values())} labelled launches ({engine}, {B} x {sec:g} s)")
    for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"]):
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
        gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
        print(f"  {k:22s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {tf:7.1f} TFLOP/s  {gb:7.0f} GB/s")
    out[name] = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in tab.items()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"train_prof_{engine}.json"), "w"), indent=1)
