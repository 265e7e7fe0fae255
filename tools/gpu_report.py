"""Stage-by-stage parity report on the GPU box (diagnostics; writes gpurun_out/report_<engine>.json).

    python tools/gpu_report.py [simt|tcgen05] [attention_variant]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tscnet_oracle as O  # noqa: E402
import synth as weights  # noqa: E402  # noqa: E402
import se_b200  # noqa: E402


def rel_max(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    engine = sys.argv[1] if len(sys.argv) > 1 else "simt"
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 16000
    rep = {"engine": engine, "attention_variant": variant, "L": L}
    sd = weights.synth_state_dict(0)
    m = se_b200.TSCNet()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.engine = engine
    m.attention_variant = variant
    enh = se_b200.EnhancerB200(m)
    noisy, clean = weights.synth_wave(2, L, seed=7, kind="speech")
    so, sg = {}, {}
    t = time.time()
    with torch.no_grad():
        yo = O.predict(noisy, sd, chunk=8, stages=so)
    rep["oracle_s"] = time.time() - t
    try:
        yg = enh(noisy.cuda(), stages=sg).cpu()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        rep["error"] = repr(e)
        yg = None
    cl = lambda t_: t_.permute(0, 2, 3, 1)
    if "in3" in sg:
        spec = torch.view_as_real(so["spec"]).permute(0, 2, 1, 3)
        rep["spec_relmax"] = rel_max(sg["in3"].cpu()[..., 1:3], spec)
        rep["spec_rel_l2"] = float((sg["in3"].cpu()[..., 1:3].double() - spec.double()).norm() / spec.double().norm())
    for k in ("encoder", "tscb1", "tscb2", "tscb3", "tscb4", "complex"):
        if k in sg:
            rep[k] = rel_max(sg[k].cpu(), cl(so[k]))
    if "mask" in sg and sg["mask"] is not None:
        rep["mask"] = rel_max(sg["mask"].cpu(), so["mask"][:, 0])
    if yg is not None:
        rep["wave_relmax"] = rel_max(yg, yo)
        rep["sisdr_delta_db"] = float((O.si_sdr(yg, clean) - O.si_sdr(yo, clean)).abs().max())
        # quick timing
        x = noisy.cuda()
        for _ in range(2):
            enh(x)
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(5):
            enh(x)
        torch.cuda.synchronize()
        rep["gpu_ms_per_call"] = (time.time() - t) / 5 * 1e3
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"report_{engine}_{variant}.json"), "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
