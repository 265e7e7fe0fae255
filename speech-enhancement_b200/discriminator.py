"""Metric discriminator of the GAN training step (/root/reference/models/discriminator.py:35-62), stock PyTorch.

SURVEY section 2 keeps this component out of kernel scope (181,650 parameters, 0.21 GFLOP per clip: 0.15 % of the generator) -- it is here
so that the training step of BASELINE configs[4] (generator + metric discriminator, core/function.py:256-313) can run end to end around the
CUDA generator: same constructor, same ``state_dict`` keys (``layers.<i>.*`` with spectral-norm ``weight_orig`` / ``weight_u`` / ``weight_v``),
same ``forward(x, y)`` on two magnitude spectrograms (B, 1, F, T)."""
from __future__ import annotations

import torch
import torch.nn as nn


class LearnableSigmoid(nn.Module):
    """beta * sigmoid(slope * x) with one learnable slope per feature (discriminator.py:7-15)"""

    def __init__(self, in_features: int, beta: float = 1.0):
        super().__init__()
        self.beta = beta
        self.slope = nn.Parameter(torch.ones(in_features))

    def forward(self, x):
        return self.beta * torch.sigmoid(self.slope * x)


class Discriminator(nn.Module):
    def __init__(self, ndf: int, in_channel: int = 2):
        super().__init__()
        sn = nn.utils.spectral_norm
        layers, c_in = [], in_channel
        for mult in (1, 2, 4, 8):                      # four stride-2 4x4 convolutions, each followed by InstanceNorm2d(affine) + PReLU
            layers += [sn(nn.Conv2d(c_in, ndf * mult, (4, 4), (2, 2), (1, 1), bias=False)), nn.InstanceNorm2d(ndf * mult, affine=True), nn.PReLU(ndf * mult)]
            c_in = ndf * mult
        layers += [nn.AdaptiveMaxPool2d(1), nn.Flatten(), sn(nn.Linear(ndf * 8, ndf * 4)), nn.Dropout(0.3), nn.PReLU(ndf * 4), sn(nn.Linear(ndf * 4, 1)),
                   LearnableSigmoid(1)]
        self.layers = nn.Sequential(*layers)

    def forward(self, x, y):
        return self.layers(torch.cat([x, y], dim=1))
