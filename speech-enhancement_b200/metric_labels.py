"""PESQ label pipeline of the metric discriminator, host side.

The reference scores every (clean, estimate) pair of a batch with the `pesq` C extension through a joblib process pool, synchronously, up to three
times per training step (/root/reference/models/discriminator.py:17-32 `pesq_loss` / `batch_pesq`, called at core/function.py:287,293,300 after a
blocking `.cpu().numpy()` of the waveforms).  On a B200 the generator step takes tens of milliseconds, so those synchronous CPU batches -- not the
discriminator's 0.21 GFLOP -- are what a real training run waits for.  This module keeps the reference's arithmetic and failure rule
(score -1 when the scorer raises; label = (score - 1) / 3.5) and changes the plumbing:

* `MetricLabelPipeline.submit(clean, est)` copies the waveforms device -> pinned host memory on a side stream (the compute stream is never blocked)
  and hands the pairs to a worker pool; it returns at once;
* `MetricLabelPipeline.result(handle)` blocks only for what is still outstanding and returns the label tensor on the requested device;
* `batch_pesq(clean_list, noisy_list)` is the reference's synchronous call on top of the same pool (drop-in for models/discriminator.py:26-32).

Workers are threads by default; `backend="process"` uses a spawned process pool like the reference's joblib workers (the scorer must then be
picklable): a numpy scorer in threads competes with the training loop for the GIL while it launches ~1000 kernels per step (measured: +12 ms on a
78 ms step with threads, +1 ms with processes).  Pinned staging buffers are recycled: a fresh cudaHostAlloc per batch costs milliseconds.

The scorer is a plain callable `score_fn(sr, clean_1d, est_1d) -> float`.  By default it is `pesq.pesq(sr, c, n, 'wb')`; the `pesq` wheel is not part
of this image, so constructing a pipeline without a scorer raises ImportError there instead of producing made-up labels.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import queue
import threading
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

ScoreFn = Callable[[int, np.ndarray, np.ndarray], float]


def _default_score_fn() -> ScoreFn:
    try:
        from pesq import pesq  # type: ignore
    except Exception as e:  # noqa: BLE001
        raise ImportError("metric_labels: the `pesq` package is not installed; pass score_fn=callable(sr, clean, est) -> float") from e
    return lambda sr, c, n: pesq(sr, c, n, "wb")


def _guarded(score_fn: ScoreFn, sr: int, c: np.ndarray, n: np.ndarray) -> float:
    """pesq_loss (discriminator.py:17-23): any exception of the scorer (silent crops make PESQ raise) becomes the score -1"""
    try:
        return float(score_fn(sr, c, n))
    except Exception:  # noqa: BLE001
        return -1.0


class _Handle:
    __slots__ = ("futures", "event", "host", "n", "pinned")

    def __init__(self):
        self.futures: List[cf.Future] = []
        self.event: Optional[torch.cuda.Event] = None
        self.host = None
        self.n = 0
        self.pinned: List[torch.Tensor] = []


def log_spectral_score(sr: int, c: np.ndarray, n: np.ndarray) -> float:
    """A stand-in scorer with PESQ's range for machines without the `pesq` wheel (bench.py --config 4): log-spectral distance over 32 ms frames
    mapped onto 1 .. 4.5.  It is NOT PESQ; it exists so that the pipeline can be exercised and timed end to end."""
    nf = (len(c) - 512) // 256 + 1
    if nf < 1 or float(np.abs(c).max()) == 0.0:
        raise ValueError("silent or too short reference")
    idx = np.arange(512)[None, :] + 256 * np.arange(nf)[:, None]
    w = np.hanning(512).astype(np.float32)
    C, N = np.abs(np.fft.rfft(c[idx] * w)) ** 2, np.abs(np.fft.rfft(n[idx] * w)) ** 2
    lsd = float(np.mean(np.sqrt(np.mean((10 * np.log10(C + 1e-8) - 10 * np.log10(N + 1e-8)) ** 2, axis=1))))
    return 4.5 - 3.5 * min(lsd / 20.0, 1.0)


class MetricLabelPipeline:
    """Asynchronous `batch_pesq`: submit() right after the generator forward, result() where the discriminator loss needs the labels."""

    def __init__(self, score_fn: Optional[ScoreFn] = None, sr: int = 16000, workers: Optional[int] = None, backend: str = "thread"):
        self.score_fn = score_fn if score_fn is not None else _default_score_fn()
        self.sr = int(sr)
        self.workers = int(workers) if workers else max(1, (os.cpu_count() or 2) - 1)
        if backend not in ("thread", "process"):
            raise ValueError("metric_labels: backend must be 'thread' or 'process'")
        self.backend = backend
        if backend == "process":
            import multiprocessing as mp
            self._pool = cf.ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context("spawn"))
        else:
            self._pool = cf.ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="seb200-metric")
        # ONE dispatcher thread sleeps on each batch's copy event (a blocking-sync event: no spinning core, never the training loop) and then hands
        # the pairs to the pool -- a dozen threads spinning in cudaEventSynchronize took cores from the thread that launches the step's kernels
        self._q: "queue.Queue" = queue.Queue()
        self._dispatcher = threading.Thread(target=self._dispatch, name="seb200-metric-dispatch", daemon=True)
        self._dispatcher.start()
        self._copy_stream = None
        self._lock = threading.Lock()
        self._free = {}                     # shape -> pinned buffers ready for reuse
        self._label_ring: List[torch.Tensor] = []
        self._label_next = 0

    def _dispatch(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            ev, hc, he, outs = item
            try:
                if ev is not None:
                    ev.synchronize()
                for b, out in enumerate(outs):
                    inner = self._pool.submit(_guarded, self.score_fn, self.sr, hc[b].numpy(), he[b].numpy())
                    inner.add_done_callback(lambda f, out=out: out.set_exception(f.exception()) if f.exception() else out.set_result(f.result()))
            except Exception as e:  # noqa: BLE001
                for out in outs:
                    if not out.done():
                        out.set_exception(e)

    # ------------------------------------------------------------------------------------------------------------------
    def _stage(self, x: torch.Tensor, h: _Handle) -> torch.Tensor:
        """detach + copy to host without blocking the caller's stream: CUDA tensors go through pinned memory on a side stream"""
        x = x.detach()
        if not x.is_cuda:
            return x.to(torch.float32).contiguous()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=x.device)
        with self._lock:
            free = self._free.get(tuple(x.shape))
            host = free.pop() if free else None
        if host is None:
            host = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
        h.pinned.append(host)
        snap = x.to(torch.float32, copy=True)      # private device snapshot on the caller's stream (0.5 MB per batch): later in-place writes to x cannot race the copy
        self._copy_stream.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(self._copy_stream):
            host.copy_(snap, non_blocking=True)
            snap.record_stream(self._copy_stream)
        return host

    def submit(self, clean: torch.Tensor, est: torch.Tensor) -> _Handle:
        """clean, est: [B, L] waveforms (any device); est may be shorter than clean (the reference crops clean to est's length, function.py:283-285)"""
        if clean.dim() != 2 or est.dim() != 2 or clean.shape[0] != est.shape[0]:
            raise ValueError("metric_labels.submit: clean and est must be [B, L] with the same B")
        h = _Handle()
        length = min(clean.shape[-1], est.shape[-1])
        hc, he = self._stage(clean[:, :length], h), self._stage(est[:, :length], h)
        h.host, h.n = (hc, he), clean.shape[0]
        if clean.is_cuda or est.is_cuda:
            h.event = torch.cuda.Event(blocking=True)
            h.event.record(self._copy_stream)
        h.futures = [cf.Future() for _ in range(h.n)]
        self._q.put((h.event, hc, he, h.futures))
        return h

    def result(self, h: _Handle, device=None) -> torch.Tensor:
        """labels (score - 1) / 3.5 as float32 [B] (discriminator.py:29-32); a failed pair carries (-1 - 1) / 3.5 like the reference"""
        scores = np.array([f.result() for f in h.futures], dtype=np.float64)
        if h.pinned:
            with self._lock:
                for buf in h.pinned:
                    self._free.setdefault(tuple(buf.shape), []).append(buf)
            h.pinned = []
        labels = torch.from_numpy(((scores - 1.0) / 3.5).astype(np.float32))
        if device is None or torch.device(device).type != "cuda":
            return labels.to(device) if device is not None else labels
        # pinned ring + asynchronous copy: a pageable H2D copy would block the host until the stream has drained (three sync points per step)
        if not self._label_ring or self._label_ring[0].numel() < labels.numel():
            self._label_ring = [torch.empty(max(labels.numel(), 64), dtype=torch.float32, pin_memory=True) for _ in range(16)]
        buf = self._label_ring[self._label_next % 16][:labels.numel()]
        self._label_next += 1
        buf.copy_(labels)
        return buf.to(device, non_blocking=True)

    def failed(self, h: _Handle) -> np.ndarray:
        """mask of the pairs whose scorer raised (the reference keeps them; a caller may want to drop the batch)"""
        return np.array([f.result() == -1.0 for f in h.futures])

    def warm_up(self):
        """start the worker processes (spawn + imports take seconds) before the first timed step"""
        if self.backend == "process":
            z = np.zeros(1024, dtype=np.float32)
            for f in [self._pool.submit(_guarded, self.score_fn, self.sr, z, z) for _ in range(self.workers)]:
                f.result()

    def close(self):
        self._q.put(None)
        self._dispatcher.join()
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


_shared: Optional[MetricLabelPipeline] = None


def batch_pesq(clean: Sequence[np.ndarray], noisy: Sequence[np.ndarray], score_fn: Optional[ScoreFn] = None, device="cuda") -> torch.Tensor:
    """The reference's synchronous call (models/discriminator.py:26-32): lists of 1-D numpy waveforms in, label tensor on `device` out."""
    global _shared
    if _shared is None or (score_fn is not None and _shared.score_fn is not score_fn):
        _shared = MetricLabelPipeline(score_fn)
    c = torch.from_numpy(np.stack([np.asarray(x, dtype=np.float32) for x in clean]))
    n = torch.from_numpy(np.stack([np.asarray(x, dtype=np.float32) for x in noisy]))
    return _shared.result(_shared.submit(c, n), device=device)
