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
        is_bn = ".conv.1." in k or ".conv.4." in k or ".bn1." in k or ".bn2." in k or k.startswith("bn1.") or \
            ".downsample.1." in k
        if is_bn and not k.endswith(("running_mean", "running_var", "num_batches_tracked")):
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


PROGRESSIVE_CASE = dict(B=1, H=128, W=256, seed=2468)


def seeded_slices(B, H, W, seed):
    """(B,5,H,W) correlated slices: a smooth drift over iid noise, each slice z-scored (ModelDataGenerator.py:73-75)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, 1, H, W, generator=g)
    drift = torch.randn(B, 1, H, W, generator=g)
    t = torch.linspace(-1.0, 1.0, 5).view(1, 5, 1, 1)
    vol = base + 0.5 * t * drift + 0.2 * torch.randn(B, 5, H, W, generator=g)
    return (vol - vol.mean(dim=(-2, -1), keepdim=True)) / (vol.std(dim=(-2, -1), keepdim=True) + 1e-6)


DEEPCNN_CASE = dict(B=1, H=32, W=48, seed=1357)


FASTDDPM_CASE = dict(B=2, H=64, W=128, seed=9753, noise_seed=77, t=(3, 8), init_seed=11)


def fastddpm_inputs(case=None):
    """cond (B,2,H,W) = (pre, post), target (B,1,H,W), t (B,), noise (B,1,H,W) drawn the way the reference's
    FastDDPM.forward / .sample draw it: first use of the global generator after torch.manual_seed(noise_seed)."""
    c = case or FASTDDPM_CASE
    sl = seeded_slices(c["B"], c["H"], c["W"], c["seed"])
    cond = sl[:, [0, 2]].contiguous()
    target = sl[:, 1:2].contiguous()
    t = torch.tensor(c["t"], dtype=torch.long)
    torch.manual_seed(c["noise_seed"])
    noise = torch.randn_like(target)
    return cond, target, t, noise


def fastddpm_state_dict(ctor, case=None):
    c = case or FASTDDPM_CASE
    torch.manual_seed(c["init_seed"])
    model = ctor(T=10, device="cpu")
    return {k: v.detach().clone() for k, v in model.state_dict().items()}
