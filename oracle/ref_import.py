"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

The reference's hot-path code only needs torch + einops, but its modules import
five packages that are absent in this image (timm, pesq, termcolor, librosa,
yacs) at file scope.  None of them is touched on the generator path, so they
are replaced by inert stubs before import.  Used by
``oracle/make_golden.py`` and by the tests that validate the oracle against the
live reference; it is never available on the GPU box and nothing there needs it.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("SE_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "generator.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    class AverageMeter:  # timm.utils.AverageMeter -- only used by the training loops
        pass

    def _pesq(*a, **k):
        raise RuntimeError("pesq is not installed (stub)")

    class CfgNode(dict):  # yacs.config.CfgNode -- config/default.py only
        def __init__(self, *a, **k):
            super().__init__(*a, **k)

        __getattr__ = dict.get

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            return CfgNode(self)

        def defrost(self):
            pass

        def freeze(self):
            pass

        def merge_from_file(self, *_):
            pass

        def merge_from_list(self, *_):
            pass

    timm = _stub("timm")
    timm.utils = _stub("timm.utils", AverageMeter=AverageMeter)
    _stub("pesq", pesq=_pesq)
    _stub("termcolor", colored=lambda s, *a, **k: s)
    _stub("librosa")
    yacs = _stub("yacs")
    yacs.config = _stub("yacs.config", CfgNode=CfgNode)


_cache = {}


def load():
    """Returns a namespace with the reference's TSCNet, predict, compressed_stft,
    uncompressed_istft and kaiming_init, imported from REF_ROOT unmodified."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise ImportError(f"reference not found under {REF_ROOT}")
    install_stubs()
    sys.path.insert(0, REF_ROOT)
    # the image's HuggingFace `datasets` package shadows the reference's datasets/ dir;
    # inference_gan imports neither for the generator path, but guard anyway
    try:
        from models.generator import TSCNet
        import models.generator as gen_mod
        import models.conformer as conf_mod
        from core.function import compressed_stft, uncompressed_istft, batch_stft, normalize_batch
        from utils.utils import kaiming_init
        try:
            from models.tsc_diffusion import TSCNet as DiffusionTSCNet
        except Exception:  # pragma: no cover
            DiffusionTSCNet = None
        try:
            from inference_diffuse import predict_tsc, inference_schedule
        except Exception:  # pragma: no cover
            predict_tsc = inference_schedule = None
        try:
            from inference_gan import predict
        except Exception:  # pragma: no cover - depends on what else the script imports
            predict = None
    finally:
        sys.path.remove(REF_ROOT)
    ns = types.SimpleNamespace(TSCNet=TSCNet, DiffusionTSCNet=DiffusionTSCNet, predict_tsc=predict_tsc,
                               inference_schedule=inference_schedule, predict=predict, compressed_stft=compressed_stft,
                               uncompressed_istft=uncompressed_istft, kaiming_init=kaiming_init,
                               batch_stft=batch_stft, normalize_batch=normalize_batch,
                               generator=gen_mod, conformer=conf_mod,
                               config=types.SimpleNamespace(N_FFT=400, HOP_SAMPLES=100))
    _cache["ns"] = ns
    return ns
