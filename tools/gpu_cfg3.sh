#!/bin/bash
# BASELINE configs[2]: 16 x 30 s utterances on one B200 (T = 4801, +-512 clamp active) and configs[0] latency (1 x 2 s)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 --batch 16 --clip-seconds 30 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "cfg3 rc=$?"; tail -c 600 gpurun_out/bench_cfg3.err
timeout 600 python bench.py --steps 20 --warmup 5 --batch 1 --clip-seconds 2 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
echo "cfg1 rc=$?"
python - <<'PY'
import json
for f in ("cfg3","cfg1"):
    try:
        d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, 'audio-s/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v for k,v in list(d['kernel_shares'].items())[:4]})
    except Exception as e: print(f, 'failed', e)
PY
timeout 900 python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import se_b200
from oracle import tscnet_oracle as O, weights
sd = weights.synth_state_dict(0)
m = se_b200.TSCNet(); m.load_state_dict(sd); m = m.cuda().eval()
noisy, clean = weights.synth_wave(1, 120000, seed=3, kind="speech")
t=time.time()
with torch.no_grad(): yo = O.predict(noisy, sd, chunk=4)
print('oracle 12 s clip:', round(time.time()-t,1), 's')
yg = se_b200.EnhancerB200(m)(noisy.cuda()).cpu()
print('T=1201 wave max-abs/peak', float((yg-yo).abs().max()/yo.abs().max()), 'si-sdr delta dB', float((O.si_sdr(yg, clean)-O.si_sdr(yo, clean)).abs().max()))
PY
