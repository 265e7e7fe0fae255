"""Randomised end-to-end parity sweep (GPU box): random batch sizes, clip lengths (including lengths that take the wrap-pad branch of predict and are no
multiple of the hop), signal kinds and weight seeds; EnhancerB200 (default tcgen05 configuration) against the oracle port on the host.
usage: fuzz_e2e.py [cases] [seed]"""
import os, sys, random, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import se_b200, synth
from oracle import tscnet_oracle as O
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
torch.set_num_threads(max(1, (os.cpu_count() or 4) - 2))
worst = 0.0
models = {}
for c in range(cases):
    B = rng.choice([1, 1, 2, 3, 5])
    L = rng.choice([rng.randint(201, 1200), rng.randint(1200, 9000), rng.randint(9000, 26000), 100 * rng.randint(20, 160), 6400 * rng.randint(1, 4)])
    kind = rng.choice(["speech", "noise"])
    wseed = rng.choice([0, 1, 2])
    if wseed not in models:
        sd = synth.synth_state_dict(wseed)
        m = se_b200.TSCNet(num_channel=64, num_features=201)
        m.load_state_dict(sd)
        models[wseed] = (se_b200.EnhancerB200(m.to("cuda").eval()), sd)
    enh, sd = models[wseed]
    noisy, _ = synth.synth_wave(B, L, seed=1000 + c, kind=kind)
    t0 = time.time()
    with torch.no_grad():
        ref = O.predict(noisy, sd)
    y = enh(noisy.to("cuda")).cpu()
    err = float((y - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    print(f"case {c:2d}: B={B} L={L:6d} {kind:6s} weights {wseed}: max-abs/peak {err:.2e}  (oracle {time.time() - t0:.1f} s)", flush=True)
    assert y.shape == ref.shape and err < 1e-3, (B, L, kind, wseed, err)
print(f"{cases} cases, worst {worst:.2e} (tolerance 1e-3)")
