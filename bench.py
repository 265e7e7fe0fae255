#!/usr/bin/env python
"""bench.py -- audio-seconds enhanced per second on the generator hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W      # CPU arm: the oracle port of the reference path

A step = one pass of the hot path (RMS-normalise -> compressed STFT -> TSCNet -> decompress + iSTFT) over one batch
of synthetic utterances: BASELINE.json configs[1], 64 x 4 s at 16 kHz per GPU.  Multi-GPU is pure batch sharding
(every rank enhances its own 64 clips; no data-path collective) => weak scaling.  `value` is timed with CUDA events
with the batch already resident in HBM; `e2e` goes through the public API from pinned host buffers with the H2D and
D2H copies inside the timed region.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
METRIC = "audio_seconds_enhanced_per_second"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[i]: 1 = 64 x 4 s per GPU (default, the metric's configuration), 2 = 16 x 30 s per GPU, "
                         "3 = 4096 x 4 s clips in total sharded across the ranks (job mode), 4 = GAN training step, 4 x 2 s per GPU (--train)")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU per step (default: what --config names)")
    ap.add_argument("--clip-seconds", type=float, default=None)
    ap.add_argument("--engine", default=None, help="tcgen05 (default) | simt")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--disc-ddp", action="store_true", help="--config 4, N > 1: keep the metric discriminator under torch's DistributedDataParallel (main_gan.py:168-171) instead of one "
                    "flat all-reduce of its 0.73 MB of gradients after its own backward (A/B of the wrapper's cost: buffer broadcasts + bucket hooks on four forward calls per step)")
    ap.add_argument("--fixed-labels", action="store_true", help="--config 4: fixed label tensors instead of the metric-label pipeline (A/B of the pipeline's cost)")
    ap.add_argument("--train", action="store_true", help="BASELINE configs[4]: generator + metric discriminator training step (same as --config 4)")
    ap.add_argument("--graphs", action="store_true", help="replay a captured CUDA graph per step (small-batch latency mode)")
    ap.add_argument("--cpu-sample-clips", type=int, default=4, help="clips of the batch the CPU legs time per step (4 x 4 s ~ 8 s of CPU work on 16 cores); "
                                                                     "the cpu_baseline leg and --impl reference use the same sample")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true", help="skip the informational PyTorch-eager leg on the same GPU")
    ap.add_argument("--total-clips", type=int, default=0,
                    help="BASELINE configs[3]: enhance this many clips in total (4096), sharded across the ranks in micro-batches of "
                         "--batch, host buffers in and out; strong scaling.  One step = the whole job.")
    args = ap.parse_args()
    if args.train:
        args.config = 4
    preset = {1: (64, 4.0, 0), 2: (16, 30.0, 0), 3: (64, 4.0, 4096), 4: (4, 2.0, 0)}[args.config]
    if args.batch is None:
        args.batch = preset[0]
    if args.clip_seconds is None:
        args.clip_seconds = preset[1]
    if args.total_clips == 0:
        args.total_clips = preset[2]
    return args


def workload_config(args, world: int):
    """the `config` object of the JSON line: a function of the command line only, so the product arm and the reference arm print the same dict"""
    B, sec = args.batch, args.clip_seconds
    named = {(64, 4.0): "BASELINE configs[1]", (16, 30.0): "BASELINE configs[2]"}.get((B, sec), "not a BASELINE configuration")
    return {"workload": f"generator hot path (RMS-norm, compressed STFT, TSCNet, iSTFT) on {B} x {sec:g} s 16 kHz utterances per GPU ({named})",
            "batch_per_gpu": B, "clip_seconds": sec, "frames": int(sec * SR) // 100 + 1, "parallelism": f"batch-shard x{world}",
            "gemm_engine": args.engine or "tcgen05", "dft_engine": "tcgen05", "cuda_graph": bool(args.graphs),
            "l2": "per-step working set (tens of GB of activations) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


# labels of the GEMM-engine / attention launches (tensor-bound); everything else is HBM-bound
TENSOR_LABELS = {"dconv", "conv2", "subpixel", "ffn1", "ffn2", "qkv", "attn_out", "pw1_glu", "pw2", "attention", "stft", "idft"}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_throughput(clips: int, clip_s: float, warmup: int, steps: int):
    """Times the oracle port of the reference path (oracle/tscnet_oracle.predict) on the host cores."""
    from oracle import tscnet_oracle as O
    import synth as weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = weights.synth_state_dict(0)
    noisy, _ = weights.synth_wave(clips, int(clip_s * SR), seed=1234, kind="speech")
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.predict(noisy, sd, chunk=16)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return clips * clip_s / t, t, cores


def cpu_sample_clips(args):
    """clips per CPU step: bounded so that one step is ~10 s of host work whatever the clip length (30 s clips: one clip)"""
    return max(1, min(args.cpu_sample_clips, int(16.0 / args.clip_seconds) or 1))


def run_reference(args):
    """the reference path on the host cores (oracle port, all threads) on the SAME bounded sample the product arm's cpu_baseline leg times"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    clips = cpu_sample_clips(args)
    val, t, cores = cpu_oracle_throughput(clips, args.clip_seconds, min(args.warmup, 1), max(1, min(args.steps, 8)))
    sample = (f"{clips} x {args.clip_seconds:g} s clips per step (of the {args.batch}-clip batch), fp32, torch CPU, all host threads; "
              f"{min(args.warmup, 1)} warm-up + {max(1, min(args.steps, 8))} timed passes (bounded: the requested {args.steps} steps of ~{t:.0f} s each would not end within minutes)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_eager_throughput(dev, clip_s: float, clips: int = 2, reps: int = 3):
    """Informational (SURVEY section 0 calls it the real bar): the reference's algorithm in PyTorch EAGER on the same B200 -- the oracle port's
    functions under torch.cuda (stock ATen / cuDNN / cuBLAS / cuFFT kernels, TF32 off), at a batch whose materialised (S, 4, n, n)
    score tensors fit.  The unmodified reference cannot travel to the GPU box; the port issues the same stock ops."""
    from oracle import tscnet_oracle as O
    import synth as weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.to(dev) for k, v in weights.synth_state_dict(0).items()}
    noisy, _ = weights.synth_wave(clips, int(clip_s * SR), seed=1234, kind="speech")
    noisy = noisy.to(dev)
    chunk = 64 if clip_s <= 4.0 else 4
    with torch.no_grad():
        O.predict(noisy, sd, chunk=chunk)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            O.predict(noisy, sd, chunk=chunk)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"value": clips * clip_s / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms, "kind": "port under torch.cuda (PyTorch eager, fp32, TF32 off)",
            "sample": f"{clips} x {clip_s:g} s clips per step, attention in chunks of {chunk} sequences"}


# ------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import se_b200
    from se_b200 import ops
    import synth as weights

    model = se_b200.TSCNet(num_channel=64, num_features=201)
    model.load_state_dict(weights.synth_state_dict(0))
    model = model.to(dev).eval()
    if args.engine:
        model.engine = args.engine
    enh = se_b200.EnhancerB200(model, use_cuda_graph=args.graphs)

    B, L = args.batch, int(args.clip_seconds * SR)
    noisy_host, _ = weights.synth_wave(B, L, seed=1234 + rank, kind="speech")
    noisy_host = noisy_host.pin_memory()
    out_host = torch.empty(B, L, dtype=torch.float32).pin_memory()
    noisy = noisy_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also builds packed weights / workspaces) + one fully instrumented step to find the dominant kernel
    for _ in range(max(args.warmup, 3)):
        enh(noisy)
    enh_eager = se_b200.EnhancerB200(model) if args.graphs else enh      # graph replays bypass the per-launch hooks
    with ops.profile() as prof:
        enh_eager(noisy)
    table = prof.summary()
    step_ms_prof = sum(v["ms"] for v in table.values())
    dominant = max(table, key=lambda k: table[k]["ms"])

    # ---- timed region: K steps, device-resident input, CUDA events; the dominant kernel's launches are event-bracketed
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = se_b200._lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.profile(only={dominant}) as dprof:
        e0.record()
        for _ in range(args.steps):
            enh(noisy)
        e1.record()
    barrier()
    launches = se_b200._lib.launch_count() - launches0
    if args.graphs:
        launches = enh.graph_kernel_nodes * args.steps                   # replayed kernel nodes (counted at capture)
    dev_ms = e0.elapsed_time(e1)
    dom = dprof.summary().get(dominant) or table[dominant]      # under --graphs: the eager profiling pass

    # ---- end to end: pinned host -> device -> enhance -> pinned host, copies inside the timed region
    for _ in range(2):
        out_host.copy_(enh(noisy_host.to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        x = noisy_host.to(dev, non_blocking=True)
        out_host.copy_(enh(x), non_blocking=True)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        audio_s = world * B * args.clip_seconds * args.steps
        pk = peaks()
        is_tensor = dominant in TENSOR_LABELS
        per_launch_ms = dom["ms"] / dom["launches"]
        if is_tensor:
            achieved = dom["flops"] / dom["launches"] / (per_launch_ms * 1e-3) / 1e12
            peak, unit, bound = pk["tflops"], "TFLOP/s", "tensor"
        else:
            achieved = dom["bytes"] / dom["launches"] / (per_launch_ms * 1e-3) / 1e9
            peak, unit, bound = pk["hbm_gbs"], "GB/s", "hbm"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dominant)
        shares = {k: round(v["ms"] / step_ms_prof, 4) for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])}
        secondary = None
        if dominant == "attention":
            # the attention core is additionally capped by the MUFU exp rate (one ex2 per (query, key, head) = flops / 96);
            # 16 ex2 / clk / SM on sm_100, at the SM clock sampled during the timed region
            gexp = dom["flops"] / 96.0 / dom["launches"] / (per_launch_ms * 1e-3) / 1e9
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            pk_exp = n_sm * 16 * sm_mhz * 1e6 / 1e9
            secondary = {"bound": "mufu", "achieved": gexp, "peak": pk_exp, "unit": "Gexp/s", "frac": gexp / pk_exp}
        # per-kernel roofline fractions (north_star) from the fully event-bracketed pass: ALGORITHMIC flops / bytes per label as the
        # wrappers in ops.py attach them, over the label's device time; `bound` is the roofline DESIGN.md section 4 assigns to the kernel
        hbm_bound = {"qkv", "attn_out", "pw1_glu", "pw2", "merge_gate", "merge_out"}
        kernel_rooflines = {}
        for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
            if v["ms"] <= 0:
                continue
            tf, gb = v["flops"] / (v["ms"] * 1e-3) / 1e12, v["bytes"] / (v["ms"] * 1e-3) / 1e9
            tensor = (k in TENSOR_LABELS or k == "ffn_fused") and k not in hbm_bound
            kernel_rooflines[k] = {"bound": "tensor" if tensor else "hbm", "ms": round(v["ms"], 3), "launches": v["launches"],
                                   "achieved": round(tf if tensor else gb, 1), "unit": "TFLOP/s" if tensor else "GB/s",
                                   "frac": round((tf / pk["tflops"]) if tensor else (gb / pk["hbm_gbs"]), 4)}
        line = {
            "metric": METRIC, "value": audio_s / (dev_ms * 1e-3), "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "generator_fwd_ms_per_clip": dev_ms / args.steps / B,
            "clocks": clocks,
            "e2e": {"value": audio_s / (e2e_ms * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": noisy_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": int(launches),
            "roofline": {"kernel": dominant, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                         "traffic": traffic, "peak_source": pk["src"], "launches_timed": dom["launches"], "ms_per_launch": per_launch_ms,
                         "share_of_step": shares.get(dominant), "secondary": secondary},
            "kernel_shares": shares,
            "kernel_rooflines": kernel_rooflines,
        }
        if world == 1 and not args.no_gpu_eager_baseline:
            try:
                del noisy
                model._ws.clear()
                torch.cuda.empty_cache()
                line["gpu_eager_baseline"] = gpu_eager_throughput(dev, args.clip_seconds)
            except Exception as exc:  # noqa: BLE001 -- informational leg: never fail the bench line
                line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            clips = cpu_sample_clips(args)
            val, tsec, cores = cpu_oracle_throughput(clips, args.clip_seconds, 0, 1)
            line["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                    "sample": f"{clips} x {args.clip_seconds:g} s clips of the batch, one pass, oracle port (torch CPU fp32), {tsec:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_job(args):
    """BASELINE configs[3] (SURVEY 8d cfg 4): N clips in total, rank r takes the contiguous slice shard_slice(N, r, G) and enhances it
    in micro-batches from pinned host memory: H2D of micro-batch i+1 and D2H of i-1 run on copy streams under the kernels of i.
    Total work is fixed as G grows => "strong" scaling.  Time = max over ranks between two barriers (device events)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import se_b200
    import synth as weights
    model = se_b200.TSCNet(num_channel=64, num_features=201)
    model.load_state_dict(weights.synth_state_dict(0))
    model = model.to(dev).eval()
    enh = se_b200.EnhancerB200(model)
    mb, L = args.batch, int(args.clip_seconds * SR)
    sl = se_b200.shard_slice(args.total_clips, rank, world)
    mine = sl.stop - sl.start
    # the rank's clips: one synthetic micro-batch tiled over the slice (content does not change the work); outputs land in one pinned buffer
    base, _ = weights.synth_wave(mb, L, seed=1234 + rank, kind="speech")
    host_in = base.pin_memory()
    host_out = torch.empty(mine, L, dtype=torch.float32).pin_memory()
    h2d, d2h, main = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    starts = list(range(0, mine, mb))

    def job():
        nxt = None
        outs = []
        with torch.cuda.stream(h2d):
            nxt = host_in[:min(mb, mine)].to(dev, non_blocking=True); ev = torch.cuda.Event(); ev.record(h2d)
        for k, s0 in enumerate(starts):
            main.wait_event(ev)
            cur = nxt
            if k + 1 < len(starts):
                n1 = min(mb, mine - starts[k + 1])
                with torch.cuda.stream(h2d):
                    nxt = host_in[:n1].to(dev, non_blocking=True); ev = torch.cuda.Event(); ev.record(h2d)
            y = enh(cur)
            cur.record_stream(main)
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                host_out[s0:s0 + y.shape[0]].copy_(y, non_blocking=True)
                y.record_stream(d2h)
            outs.append(y)
        main.wait_stream(d2h)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        enh(host_in.to(dev))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = se_b200._lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        job()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        ms = float(ms[0])
        val = args.total_clips * args.clip_seconds * args.steps / (ms * 1e-3)
        line = {"metric": METRIC, "value": val, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"batch-sharded inference of {args.total_clips} x {args.clip_seconds:g} s utterances across {world} GPU(s) "
                                       f"(BASELINE configs[3]), micro-batches of {mb}, pinned host buffers in and out, copies overlapped",
                           "total_clips": args.total_clips, "micro_batch": mb, "parallelism": f"batch-shard x{world}",
                           "l2": "per-step working set >> 126 MB L2; no flush needed"},
                "clocks": clocks,
                "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": mine * L * 4, "d2h_bytes_per_step": mine * L * 4},
                "gpu_launches": int(se_b200._lib.launch_count() - launches0)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_METRIC = "audio_seconds_trained_per_second"
LOSS_WEIGHTS = (0.3, 0.7, 0.2, 0.05)          # config/scp.yaml:7


def train_config(args, world):
    return {"workload": f"GAN training step (generator + metric discriminator fwd/bwd, arch scp: consistency-preserving losses), {args.batch} x {args.clip_seconds:g} s "
                        f"clips per GPU, gradient all-reduce over {world} GPU(s) (BASELINE configs[4])",
            "batch_per_gpu": args.batch, "clip_seconds": args.clip_seconds, "frames": int(args.clip_seconds * SR) // 100 + 1,
            "parallelism": f"data-parallel x{world} (SyncBatchNorm statistics + one flat 7.34 MB gradient all-reduce; discriminator: " +
                           ("DistributedDataParallel)" if args.disc_ddp else "one coalesced 0.73 MB gradient all-reduce after its backward)"),
            "generator_engine": args.engine or "tcgen05_f32", "optimizer": "AdamW (fused)", "pesq_labels": ("fixed label tensors (--fixed-labels)" if args.fixed_labels else
                            "three label batches per step through se_b200.MetricLabelPipeline (side-stream D2H + worker processes, models/discriminator.py:17-32); the scorer is a "
                            "log-spectral-distance stand-in with PESQ's range because the pesq wheel is not in the image (SURVEY 8d cfg 5)"),
            "l2": "per-step working set ~10 GB of saved activations >> 126 MB L2; no flush needed"}


def _gan_step(se_b200, model, disc, opt_g, opt_d, batch, cfg, args_ns, world, mse, timers=None, labels=None, disc_sync=None):
    """one iteration of train_gan (core/function.py:206-317, arch scp) on the CUDA generator; `timers`: dict of (start, end) event lists"""
    import torch.nn.functional as F

    def tick(name):
        if timers is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        timers.setdefault(name, []).append(e)
        return e

    tick("start")
    opt_g.zero_grad(set_to_none=True)
    clean, noisy, clean_spec, noisy_spec, clean_real, clean_imag, one_labels, win = se_b200.batch_stft(batch, args_ns, cfg)
    tick("stft")
    est_real, est_imag = model(noisy_spec)
    tick("gen_fwd")
    est_real, est_imag = est_real.permute(0, 1, 3, 2), est_imag.permute(0, 1, 3, 2)
    est_complex = torch.complex(est_real, est_imag).squeeze(1)
    est_mag = est_complex.abs().unsqueeze(1)
    clean_mag = clean_spec.abs().unsqueeze(1)
    est_audio = se_b200.uncompressed_istft(est_complex, 400, 100, win)
    if labels is not None:      # the three PESQ batches of the discriminator step (:287,293,300) start now and are collected where their losses need them
        hq = (labels.submit(clean, est_audio), labels.submit(clean, clean), labels.submit(clean, noisy))
    est_prime = se_b200.compressed_stft(est_audio, 400, 100, win)                       # consistency branch (:231-254)
    with torch.no_grad():
        clean_prime_audio = se_b200.uncompressed_istft(clean_spec, 400, 100, win)
        clean_prime = se_b200.compressed_stft(clean_prime_audio, 400, 100, win)
    loss_mag = mse(est_prime.abs(), clean_prime.abs())
    time_loss = torch.mean(torch.abs(est_audio - clean_prime_audio))
    loss_ri = mse(est_prime.real, clean_prime.real) + mse(est_prime.imag, clean_prime.imag)
    gan = mse(disc(clean_mag, est_mag).flatten(), one_labels.float())
    loss = LOSS_WEIGHTS[0] * loss_ri + LOSS_WEIGHTS[1] * loss_mag + LOSS_WEIGHTS[2] * time_loss + LOSS_WEIGHTS[3] * gan
    tick("losses_fwd")
    loss.backward()
    tick("gen_bwd")
    se_b200.allreduce_gradients(model)
    tick("allreduce")
    opt_g.step()
    tick("opt_g")
    # ---- discriminator step (:279-313): L_E + L_C + L_N against the metric labels
    opt_d.zero_grad(set_to_none=True)
    if labels is not None:
        q_e, q_c, q_n = (labels.result(h, device=one_labels.device) for h in hq)
    else:
        q_e, q_c, q_n = torch.full_like(one_labels, 0.6), one_labels, torch.full_like(one_labels, 0.3)
    d_loss = mse(disc(clean_mag, est_mag.detach()).flatten(), q_e) + mse(disc(clean_mag, clean_mag).flatten(), q_c) + \
        mse(disc(clean_mag, noisy_spec.abs().unsqueeze(1)).flatten(), q_n)
    d_loss.backward()
    if disc_sync is not None:                      # data-parallel exchange of the discriminator step: one coalesced all-reduce (mean) of its gradients
        se_b200.allreduce_gradients(disc_sync)
    opt_d.step()
    tick("disc")
    return loss.detach(), d_loss.detach()


def run_train(args):
    """BASELINE configs[4]: one GAN training step per bench step, B clips of clip_seconds per GPU, weak scaling."""
    import types
    import torch.distributed as dist
    import torch.nn as nn
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --train: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import se_b200
    from se_b200.discriminator import Discriminator
    import synth as weights
    torch.manual_seed(1234)
    model = se_b200.TSCNet(num_channel=64, num_features=201)
    model.load_state_dict(weights.synth_state_dict(0))
    disc = Discriminator(16)
    if world > 1:                                   # main_gan.py:154-171
        model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model, disc = model.to(dev).train(), disc.to(dev).train()
    st = se_b200.training._state(model)
    if args.engine:
        st.engine = args.engine
    disc_run = nn.parallel.DistributedDataParallel(disc, device_ids=[local]) if (world > 1 and args.disc_ddp) else disc
    disc_sync = disc if (world > 1 and not args.disc_ddp) else None     # identical replicas (same seed, same updates): no buffer broadcast needed
    opt_g = torch.optim.AdamW(model.parameters(), lr=5e-4, fused=True)
    opt_d = torch.optim.AdamW(disc.parameters(), lr=1e-3, fused=True)
    mse = nn.MSELoss()
    B, L = args.batch, int(args.clip_seconds * SR)
    noisy, clean = weights.synth_wave(B, L, seed=1234 + rank, kind="speech")
    host = {"audio": clean.pin_memory(), "noisy": noisy.pin_memory()}
    cfg = types.SimpleNamespace(N_FFT=400, HOP_SAMPLES=100)
    args_ns = types.SimpleNamespace(gpu=local)

    labels = None
    if not args.fixed_labels:
        labels = se_b200.MetricLabelPipeline(se_b200.metric_labels.log_spectral_score, workers=min(6, max(2, (os.cpu_count() or 4) // (2 * max(world, 1)))), backend="process")
        labels.warm_up()

    def step(timers=None):
        return _gan_step(se_b200, model, disc_run, opt_g, opt_d, host, cfg, args_ns, world, mse, timers, labels, disc_sync)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    timers = {}
    step(timers)                                    # one instrumented step: phase breakdown
    torch.cuda.synchronize()
    order = ["start", "stft", "gen_fwd", "losses_fwd", "gen_bwd", "allreduce", "opt_g", "disc"]
    phases = {order[i + 1]: timers[order[i]][0].elapsed_time(timers[order[i + 1]][0]) for i in range(len(order) - 1)}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = se_b200._lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g_loss, d_loss = step()
    e1.record()
    barrier()
    launches = se_b200._lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    spread = None
    if world > 1:                                   # the replicas must still be identical: largest difference of a parameter checksum across ranks (generator, discriminator)
        cs = torch.stack([sum(p.detach().double().abs().sum() for p in m.parameters()) for m in (model, disc)])
        both = torch.cat([cs, -cs])
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        spread = [float((both[i] + both[i + 2]) / both[i].abs().clamp_min(1e-30)) for i in range(2)]
    if rank == 0:
        ms = float(ms[0])
        audio_s = world * B * args.clip_seconds * args.steps
        val = audio_s / (ms * 1e-3)
        # roofline of the step's dominant phase: generator backward, algorithmic FLOPs = 2 x forward (dgrad + wgrad of every contraction)
        T = L // 100 + 1
        fwd_flops = B * (643200.0 * T + 101.36e6 * T + 109.4e6 * T + 404.0 * T * (480640.0 + 384.0 * T))
        pk = peaks()
        bwd_tf = 2.0 * fwd_flops / (phases["gen_bwd"] * 1e-3) / 1e12
        line = {"metric": TRAIN_METRIC, "value": val, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": train_config(args, world), "clocks": clocks,
                "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 2 * B * L * 4 + 3 * B * 4, "d2h_bytes_per_step": 6 * B * L * 4,
                        "note": "every step starts from pinned host waveforms (batch_stft copies them to the device), as the reference's loader does; the three metric-label "
                                "batches copy clean / estimate / noisy waveforms back to pinned host memory on a side stream and return B labels each"},
                "gpu_launches": int(launches),
                "phases_ms": {k: round(v, 3) for k, v in phases.items()},
                "collective": {"gradient_allreduce_bytes": 1834833 * 4, "allreduce_ms": round(phases["allreduce"], 3),
                               "syncbn_allreduces_per_step": 16 if world > 1 else 0, "limiting": "latency (7.34 MB + 16 x 2 KB): NCCL launch + NVLS one-shot",
                               "discriminator": "DistributedDataParallel" if (world > 1 and args.disc_ddp) else ("one coalesced all-reduce" if world > 1 else None),
                               "replica_checksum_spread_rel": spread},
                "roofline": {"kernel": "generator backward (all kernels)", "bound": "tensor", "achieved": bwd_tf, "peak": pk["tflops"], "unit": "TFLOP/s",
                             "frac": bwd_tf / pk["tflops"], "traffic": None, "peak_source": pk["src"], "ms": phases["gen_bwd"]},
                "losses": {"generator": float(g_loss), "discriminator": float(d_loss)}}
        if world == 1 and not args.no_gpu_eager_baseline:
            try:
                line["gpu_eager_baseline"] = gpu_eager_train(dev, B, L)
            except Exception as exc:  # noqa: BLE001
                line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:200]}
        print(json.dumps(line), flush=True)
    if labels is not None:
        labels.close()
    if world > 1:
        dist.destroy_process_group()


def gpu_eager_train(dev, B, L, reps=3):
    """Informational: generator forward + backward of the reference's algorithm in PyTorch eager on the same GPU (oracle port with autograd,
    train mode, its own dropout draw), fp32, TF32 off -- the part of the step the CUDA generator replaces."""
    from oracle import tscnet_oracle as O
    import synth as weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: (v.to(dev).requires_grad_(v.is_floating_point() and "running_" not in k)) for k, v in weights.synth_state_dict(0).items()}
    noisy, _ = weights.synth_wave(B, L, seed=1234, kind="speech")
    c = torch.sqrt(L / torch.sum(noisy ** 2.0, dim=-1, keepdim=True))
    spec = O.compressed_stft((noisy * c).to(dev))
    T = spec.shape[-1]
    masks = {k: v.to(dev) for k, v in weights.masks_reference_layout(weights.dropout_masks(0, B, T, 101)).items()}

    def once():
        tr = O.TrainCtx(masks)
        fr, fi = O.tscnet_forward(spec, sd, tr=tr)
        (fr.square().mean() + fi.square().mean()).backward()
        for v in sd.values():
            v.grad = None
    once()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"generator_fwd_bwd_ms": ms, "kind": "port under torch.cuda with autograd (PyTorch eager, fp32, TF32 off)", "sample": f"{B} x {L / SR:g} s clips"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4 or args.train:
        run_train(args)
    elif args.total_clips > 0:
        run_job(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
