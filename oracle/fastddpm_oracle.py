"""TEST INFRASTRUCTURE ONLY — CPU restatement (functional, on a state_dict) of the reference Fast-DDPM registry model.

Follows /root/reference/src/ModelLoader.py: sinusoidal_timestep_embedding :475-487, FastNoiseScheduler :490-518,
DoubleConv :521-533, UNet2D.forward :561-585 (time embedding TILED over space and concatenated to the 3 image
channels, 259-channel first conv, F.max_pool2d, nearest F.interpolate, torch.cat([up, skip])), FastDDPM.forward
:595-602 (noise-prediction MSE) and FastDDPM.sample :604-636 (deterministic DDIM, clamp(-1,1)).

PINNED: oracle/make_golden.py imports the unmodified reference classes in the build container and checks this file
against them (state_dict identity, eps, loss, every gradient, a full 10-step sample) before writing
tests/golden/fastddpm_golden.npz. The reference draws its noise inside forward()/sample() (torch.randn_like /
torch.randn); the oracle takes the noise as an argument and make_golden.py reproduces the reference's draw by seeding
the global generator immediately before the call.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim=256):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


def schedule(T=10):
    """alpha_bar of the T selected steps (:497-513) and their indices into the 1000-step linear-beta chain."""
    beta = torch.linspace(1e-4, 0.02, 1000)
    alpha_bar = torch.cumprod(1.0 - beta, 0)
    boundary = 699
    late = int(T * 0.6)
    early = T - late
    idxs = torch.sort(torch.cat([torch.linspace(0, boundary, early).long(), torch.linspace(boundary, 999, late).long()]))[0]
    return alpha_bar[idxs], idxs


def _double_conv(sd, prefix, x):
    x = F.relu(F.conv2d(x, sd[f"{prefix}.block.0.weight"], sd[f"{prefix}.block.0.bias"], padding=1))
    return F.relu(F.conv2d(x, sd[f"{prefix}.block.2.weight"], sd[f"{prefix}.block.2.bias"], padding=1))


def unet2d_forward(sd, x, t, prefix="unet."):
    """x: (B,3,H,W) = [x_t, pre, post]; t: (B,) long. sd keys as in FastDDPM.state_dict() (prefix 'unet.')."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    e = timestep_embedding(t, 256)
    e = F.linear(F.relu(F.linear(e, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])), sd["time_mlp.2.weight"],
                 sd["time_mlp.2.bias"])
    e = e[:, :, None, None].repeat(1, 1, x.shape[2], x.shape[3])
    x = torch.cat([x, e], dim=1)
    c1 = _double_conv(sd, "inc", x)
    c2 = _double_conv(sd, "down1", F.max_pool2d(c1, 2))
    c3 = _double_conv(sd, "down2", F.max_pool2d(c2, 2))
    u2 = _double_conv(sd, "up2", torch.cat([F.interpolate(c3, scale_factor=2), c2], dim=1))
    u1 = _double_conv(sd, "up1", torch.cat([F.interpolate(u2, scale_factor=2), c1], dim=1))
    return F.conv2d(u1, sd["outc.weight"], sd["outc.bias"])


def q_sample(x0, t, noise, T=10):
    a_bar = schedule(T)[0][t].view(-1, 1, 1, 1)
    return torch.sqrt(a_bar) * x0 + torch.sqrt(1 - a_bar) * noise


def loss_and_grads(sd, cond, target, t, noise, T=10):
    """FastDDPM.forward (:595-602) + backward. Returns (loss, eps_pred, {name: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    x_t = q_sample(target, t, noise, T)
    eps = unet2d_forward(leaves, torch.cat([x_t, cond], dim=1), t)
    loss = F.mse_loss(eps, noise)
    loss.backward()
    return loss.detach(), eps.detach(), {k: v.grad.detach() for k, v in leaves.items()}


@torch.no_grad()
def sample(sd, cond, x_T, T=10):
    """FastDDPM.sample (:604-636) from a given initial noise x_T."""
    a_bars = schedule(T)[0]
    x = x_T.clone()
    B = cond.shape[0]
    for i in reversed(range(T)):
        t = torch.full((B,), i, dtype=torch.long)
        eps = unet2d_forward(sd, torch.cat([x, cond], 1), t)
        a_bar = a_bars[i]
        a_prev = a_bars[i - 1] if i > 0 else torch.tensor(1.0)
        x0 = (x - torch.sqrt(1 - a_bar) * eps) / torch.sqrt(a_bar)
        x = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev) * eps
    return x.clamp(-1, 1)


def clip_and_adam(sd, grads, opt_state, step, lr=2e-4, max_norm=1.0, b1=0.9, b2=0.999, eps=1e-8):
    """clip_grad_norm_(max_norm) + torch.optim.Adam step, functional. Returns the new state_dict."""
    names = [k for k in sd if k in grads]
    total = torch.sqrt(sum(grads[k].double().pow(2).sum() for k in names)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    new = dict(sd)
    for k in names:
        g = grads[k] * coef
        m, v = opt_state.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        opt_state[k] = (m, v)
        denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
        new[k] = sd[k] - (lr / (1 - b1 ** step)) * (m / denom)
    return new
