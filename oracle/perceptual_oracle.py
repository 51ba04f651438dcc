"""TEST INFRASTRUCTURE ONLY — CPU oracle of the VGG16 perceptual term (definition frozen in
multi-image-super-resolution-for-medical-images_b200/perceptual.py; PARITY UNPINNED against the reference, whose
combined-loss notebook is missing from the snapshot — README.md:82-86 is the only evidence)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

_CONV_IDX = (0, 2, 5, 7, 10, 12, 14)
_POOL_AFTER = (2, 7)


def vgg_features(sd, x):
    """torchvision VGG16 features[:16] (conv1_1 .. relu3_3) on a single-channel input replicated to 3 channels."""
    x = x.repeat(1, 3, 1, 1)
    for i in _CONV_IDX:
        x = torch.relu(F.conv2d(x, sd[f"features.{i}.weight"].to(x.dtype), sd[f"features.{i}.bias"].to(x.dtype), padding=1))
        if i in _POOL_AFTER:
            x = F.max_pool2d(x, 2, 2)
    return x


def perceptual_loss(sd, pred, target, weight=0.01):
    return weight * F.mse_loss(vgg_features(sd, pred), vgg_features(sd, target))
