"""Two-GPU check of the training step's data-parallel path (SURVEY 8e training; main_gan.py:154-171), launched by torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/train_ddp_check.py

Rank r trains on utterances [2r, 2r + 2) of a 4-utterance batch with SyncBatchNorm (global batch statistics through the all-reduce of the
local sums) and the flat gradient all-reduce; rank 0 also runs the whole batch alone with plain BatchNorm1d.  The summed gradients of the
two ranks must equal the single-process gradients of the full batch, and the running statistics must agree."""
import os, sys, json
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se_b200, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
engine = sys.argv[1] if len(sys.argv) > 1 else "simt"
B, L = 2 * world, 8000
sd = synth.synth_state_dict(0)
noisy, _ = synth.synth_wave(B, L, 77, "speech")
spec_all = se_b200.compressed_stft((noisy * torch.sqrt(L / noisy.pow(2).sum(-1, keepdim=True))).to(dev))
Tn = spec_all.shape[-1]
masks = synth.dropout_masks(5, B, Tn, 101)
gr, gi = synth.cotangents(3, B, Tn)
gr, gi = gr.to(dev), gi.to(dev)


def run(model, sl):
    st = se_b200.training._state(model)
    st.engine = engine
    st.injected_masks = {k: v[sl].contiguous() for k, v in masks.items()}
    fr, fi = model(spec_all[sl].contiguous())
    ((fr * gr[sl]).sum() + (fi * gi[sl]).sum()).backward()
    return st


m = se_b200.TSCNet(64, 201)
m.load_state_dict(sd)
m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m).to(dev).train()
sl = slice(2 * rank, 2 * rank + 2)
st = run(m, sl)
assert st.grad_buffer() is not None
se_b200.allreduce_gradients(m, average=False)           # sum over ranks == gradient of the summed loss over the whole batch
torch.cuda.synchronize()
ok = True
if rank == 0:
    ref = se_b200.TSCNet(64, 201)
    ref.load_state_dict(sd)
    ref = ref.to(dev).train()
    run(ref, slice(0, B))
    worst, wk = 0.0, ""
    errs = []
    for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        if synth.has_zero_gradient(k):
            continue
        e = float((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-30))
        errs.append(e)
        if e > worst:
            worst, wk = e, k
    med = sorted(errs)[len(errs) // 2]
    bn_err = max(float((a - b).abs().max()) for (ka, a), (kb, b) in zip(m.state_dict().items(), ref.state_dict().items()) if "running_" in ka)
    res = {"engine": engine, "world": world, "median_rel_l2": med, "worst_rel_l2": worst, "worst_key": wk, "running_stats_max_abs_diff": bn_err}
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"train_ddp_check_{engine}.json"), "w"))
    # two fp32 runs of this network that differ in one ulp somewhere agree to ~1e-3 in their gradients (it amplifies rounding ~100x); the
    # running statistics and the exchange itself are exact to an ulp
    ok = med < 2e-3 and worst < 2e-2 and bn_err < 1e-5
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
