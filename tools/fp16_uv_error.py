"""VERDICT r1 item 3 experiment (CPU, oracle): the conformer conv module with the GLU output u and the depthwise output v stored in fp16
(what the CUDA path writes between pw1 -> depthwise -> pw2): whole-path waveform error against the reference goldens."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, numpy as np, synth
from oracle import tscnet_oracle as O
import torch.nn.functional as F
from conftest import load_golden
orig=O.conv_module
def conv_module_h(x, sd, p, tr=None, round_u=True, round_v=True):
    h = O._ln(x, sd, p + ".net.0").transpose(1, 2)
    h = F.conv1d(h, sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])
    a, g = h.chunk(2, dim=1)
    h = a * torch.sigmoid(g)
    if round_u: h = h.half().float()
    h = F.pad(h, (15, 15))
    h = F.conv1d(h, sd[p + ".net.4.conv.weight"], sd[p + ".net.4.conv.bias"], groups=h.shape[1])
    h = F.batch_norm(h, sd[p + ".net.5.running_mean"], sd[p + ".net.5.running_var"], sd[p + ".net.5.weight"], sd[p + ".net.5.bias"], False, 0.0, 1e-5)
    h = O._swish(h)
    if round_v: h = h.half().float()
    h = F.conv1d(h, sd[p + ".net.7.weight"], sd[p + ".net.7.bias"])
    return h.transpose(1, 2)
for name in ["speech_b2_L8000","speech_b1_L16000"]:
    g=load_golden(name)
    sd=synth.synth_state_dict(int(g["weight_seed"]))
    noisy=torch.from_numpy(g["noisy"]); ref=torch.from_numpy(g["enhanced"])
    for ru,rv in [(True,True),(True,False),(False,True)]:
        O.conv_module=lambda x,sd,p,tr=None,ru=ru,rv=rv: conv_module_h(x,sd,p,tr,ru,rv)
        with torch.no_grad(): y=O.predict(noisy,sd)
        print(name,"u16" if ru else "u32","v16" if rv else "v32","err %.2e"%float((y-ref).abs().max()/ref.abs().max()))
    O.conv_module=orig
