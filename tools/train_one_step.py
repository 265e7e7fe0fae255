"""One generator training step (train-mode forward + backward) at BASELINE configs[4]'s per-GPU shape, for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_one_step.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se_b200, synth
engine = sys.argv[1] if len(sys.argv) > 1 else "tcgen05_f32"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B, L = 4, 32000
m = se_b200.TSCNet(64, 201)
m.load_state_dict(synth.synth_state_dict(0))
m = m.cuda().train()
se_b200.training._state(m).engine = engine
noisy, _ = synth.synth_wave(B, L, 1234, "speech")
spec = se_b200.compressed_stft(noisy.cuda())
for _ in range(warm + 1):
    fr, fi = m(spec)
    (fr.square().mean() + fi.square().mean()).backward()
    m.zero_grad()
torch.cuda.synchronize()
print("done")
