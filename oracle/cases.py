"""TEST INFRASTRUCTURE ONLY — seeded inputs / weights shared by make_golden.py, the tests, smoke() and bench.py.

Weights: `torch.manual_seed(0); UNet()` (default PyTorch init, reference unet_model.py:45-80; no saved checkpoints
exist in the reference snapshot) with the BatchNorm affine parameters perturbed by seeded noise so that gamma/beta
paths are exercised. Inputs: iid N(0,1) like the reference's own create_dummy_dataset (unet_model.py:301-310).
"""
from __future__ import annotations

import torch


def seeded_state_dict(model_ctor, seed=0, perturb_seed=7):
    torch.manual_seed(seed)
    model = model_ctor()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(perturb_seed)
    for k in sd:
        if ".conv.1." in k or ".conv.4." in k:
            if k.endswith(".weight"):
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
            elif k.endswith(".bias"):
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    return sd


def seeded_batch(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 2, H, W, generator=g)
    y = torch.randn(B, 1, H, W, generator=g)
    return x, y


TRAIN_CASE = dict(B=2, H=128, W=256, seed=1234)
EVAL_CASE = dict(B=1, H=256, W=256, seed=4321)
GRAD_HEAD = 32  # leading elements of every gradient stored verbatim in the golden file
