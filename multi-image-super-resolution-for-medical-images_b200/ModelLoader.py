"""Mirror of the reference model registry / checkpoint loader (reference: src/ModelLoader.py:642-711).

Same `load_model(model_name, device)` contract: known names, `ValueError` for unknown names,
`FileNotFoundError` for a missing checkpoint, three accepted checkpoint layouts (`generator_state_dict`,
`model_state_dict`, bare state_dict), result returned in eval mode. Every name of the reference registry is
implemented natively on the b200sr kernels: 'unet' / 'unet_combined' (the hot path), 'progressive_unet', 'deepcnn',
'unet_gan' (the generator; it is structurally a UNetStage) and 'fastddpm'. There is no fallback to torch modules.
"""
from __future__ import annotations

import os

import torch

from .unet_model import UNet, UNetBlock  # noqa: F401  (re-exported like the reference module does)
from .progressive import (GANUNetBlock, ProgressiveUNet, ProgressiveUNetBlock, UNetGenerator,  # noqa: F401
                          UNetStage)
from .deepcnn import DeepCNN, ResidualBlock  # noqa: F401
from .fastddpm import (DoubleConv, FastDDPM, FastNoiseScheduler, UNet2D,  # noqa: F401
                       sinusoidal_timestep_embedding)

_UNET_KW = {'in_channels': 2, 'out_channels': 1, 'init_features': 64}

# name -> (checkpoint filename, model class or None when out of scope, init kwargs)
CHECKPOINT_MAP = {
    'unet': ('unet_best.pt', UNet, _UNET_KW),
    'unet_combined': ('unet_combined_best.pt', UNet, _UNET_KW),
    'deepcnn': ('deepcnn_best.pt', DeepCNN, {'in_channels': 2, 'out_channels': 1, 'num_blocks': [2, 2, 2, 2],
                                             'base_features': 64}),
    'progressive_unet': ('progressive_unet_best.pt', ProgressiveUNet, {'base_features': 64}),
    'unet_gan': ('unet_gan_best.pt', UNetGenerator, {'in_channels': 2, 'out_channels': 1, 'base_features': 64}),
    'fastddpm': ('fastddpm_advanced_best.pth', FastDDPM, {'T': 10}),
}


def _default_root() -> str:
    # reference: parent of the directory holding ModelLoader.py; overridable for deployments
    return os.environ.get("B200SR_MODEL_ROOT",
                          os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_model(model_name, device='cuda', root=None, verbose=True):
    """Load the best checkpoint of `model_name` into the b200sr model; eval mode, on `device`.

    `root` (addition): directory that holds `models/` and `notebooks/`; defaults to $B200SR_MODEL_ROOT or the
    repository root, mirroring the reference's `<repo>/models` then `<repo>/notebooks` lookup (:657-682)."""
    key = str(model_name).lower()
    if key not in CHECKPOINT_MAP:
        raise ValueError(f"Unknown model: {model_name}. Choose from: {list(CHECKPOINT_MAP.keys())}")
    filename, cls, kwargs = CHECKPOINT_MAP[key]
    root = root or _default_root()
    path = os.path.join(root, 'models', filename)
    if not os.path.exists(path):
        path = os.path.join(root, 'notebooks', filename)
    if not os.path.exists(path):
        raise FileNotFoundError(f"Checkpoint not found: {path}")
    if cls is None:
        raise NotImplementedError(
            f"'{key}' is in the reference registry but outside the b200sr hot path (UNet only); "
            "load it with the reference ModelLoader")
    if cls is FastDDPM:  # the reference passes the device into the constructor (:668)
        kwargs = dict(kwargs, device=device)
    model = cls(**kwargs).to(device)
    checkpoint = torch.load(path, map_location=device)
    if isinstance(checkpoint, dict) and 'generator_state_dict' in checkpoint:
        state = checkpoint['generator_state_dict']
    elif isinstance(checkpoint, dict) and 'model_state_dict' in checkpoint:
        state = checkpoint['model_state_dict']
    else:
        state = checkpoint
    model.load_state_dict(state)
    model.eval()
    if verbose:
        print(f"Loaded {key.upper()} model from {os.path.basename(path)}")
    return model
