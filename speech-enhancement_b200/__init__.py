"""se_b200 -- B200-native (sm_100a) generator hot path of minyoungpark1/Speech-Enhancement.

Public surface (mirrors the reference's interface for this path):

* ``TSCNet(num_channel=64, num_features=201)``            models/generator.py:132-167
* ``compressed_stft`` / ``uncompressed_istft``             core/function.py:685-703
* ``normalize_batch`` / ``batch_stft``                       core/function.py:647-683 (training caller, forward DSP)
* ``EnhancerB200(model)(noisy)`` / ``.predict(numpy)``     inference_gan.py:75-100
* ``load_model(path)``                                      inference_gan.py:60-72
* ``shard_slice`` / ``enhance_sharded``                     batch sharding across ranks (inference has no collective)
* ``TSCNet.train()`` + ``forward`` + ``loss.backward()``      core/function.py:218-277: training.py; ``allreduce_gradients``
  = the DDP exchange of main_gan.py:168-171 as one flat all-reduce
* ``MetricLabelPipeline`` / ``batch_pesq``                   models/discriminator.py:17-32: the PESQ label batches of the discriminator
  step as an asynchronous host pipeline (side-stream copies + worker pool); ``discriminator.Discriminator`` is the stock-PyTorch module
* ``tsc_diffusion.TSCNet(num_channel, num_features, noise_schedule)``   models/tsc_diffusion.py:43-90
* ``diffusion.predict_tsc`` / ``DiffusionEnhancerB200``      inference_diffuse.py:231-267 (reverse process on the same kernels)

The directory is called ``speech-enhancement_b200``; import it as ``se_b200`` (the ``se_b200.py`` shim at the repo root).
"""
from .generator import TSCNet
from .dsp import batch_stft, compressed_stft, normalize_batch, uncompressed_istft
from .enhancer import EnhancerB200
from .sharding import enhance_sharded, shard_slice
from . import ops, packing, _lib, tsc_diffusion, diffusion, train_ops, training, metric_labels
from .metric_labels import MetricLabelPipeline, batch_pesq
from .training import allreduce_gradients
from .diffusion import DiffusionEnhancerB200, predict_tsc


def load_model(model_path, device="cuda"):
    """inference_gan.load_model: build TSCNet(64, 201), load ckpt['gen_state_dict'] with the 'module.' prefix stripped, eval()."""
    import torch
    model = TSCNet(num_channel=64, num_features=201).to(device)
    ckpt = torch.load(model_path, map_location=device)
    sd = ckpt["gen_state_dict"] if "gen_state_dict" in ckpt else ckpt
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    model.load_state_dict(sd)
    model.eval()
    return model


__all__ = ["TSCNet", "compressed_stft", "uncompressed_istft", "normalize_batch", "batch_stft", "EnhancerB200", "load_model", "shard_slice",
           "enhance_sharded", "ops", "packing", "tsc_diffusion", "diffusion", "DiffusionEnhancerB200", "predict_tsc", "training", "train_ops", "allreduce_gradients", "metric_labels", "MetricLabelPipeline", "batch_pesq"]
