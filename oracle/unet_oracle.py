"""TEST INFRASTRUCTURE ONLY — functional CPU (fp32/fp64) restatement of the reference UNet and its train step.

Follows, op by op, /root/reference/src/unet_model.py:
  UNetBlock (:22-36)  conv3x3(p=1,bias) -> BatchNorm2d(eps 1e-5, momentum 0.1) -> ReLU, twice
  UNet.forward (:82-118)  4 encoder blocks with MaxPool2d(2,2), bottleneck, 4 x (ConvTranspose2d(k2,s2) ->
                          cat([up, skip], dim=1) -> block), Conv2d 1x1 head
  UNetTrainer.train_epoch (:168-191)  forward, MSELoss, zero_grad, backward, Adam(lr 1e-4).step
It works directly on a state_dict (names exactly as the reference registers them), so it can be checked against
the reference modules (oracle/make_golden.py) and run on the GPU box where /root/reference does not exist.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ENCODERS = ("enc1", "enc2", "enc3", "enc4")
DECODERS = (("upconv4", "dec4"), ("upconv3", "dec3"), ("upconv2", "dec2"), ("upconv1", "dec1"))


def _block(sd, prefix, x, training, new_stats):
    for conv_i, bn_i in ((0, 1), (3, 4)):
        # conv bias is present in UNetBlock (unet_model.py:27,30) and absent in ProgressiveUNetBlock (ModelLoader.py:38,41)
        x = F.conv2d(x, sd[f"{prefix}.conv.{conv_i}.weight"], sd.get(f"{prefix}.conv.{conv_i}.bias"), padding=1)
        g, b = sd[f"{prefix}.conv.{bn_i}.weight"], sd[f"{prefix}.conv.{bn_i}.bias"]
        rm, rv = sd[f"{prefix}.conv.{bn_i}.running_mean"], sd[f"{prefix}.conv.{bn_i}.running_var"]
        if training:
            # F.batch_norm updates the running statistics in place: give it copies and hand them back
            rm_new, rv_new = rm.detach().clone(), rv.detach().clone()
            x = F.batch_norm(x, rm_new, rv_new, g, b, training=True, momentum=BN_MOMENTUM, eps=BN_EPS)
            if new_stats is not None:
                new_stats[f"{prefix}.conv.{bn_i}.running_mean"] = rm_new
                new_stats[f"{prefix}.conv.{bn_i}.running_var"] = rv_new
        else:
            x = F.batch_norm(x, rm, rv, g, b, training=False, momentum=BN_MOMENTUM, eps=BN_EPS)
        x = torch.relu(x)
    return x


def unet_forward(sd, x, training=False, new_stats=None, prefix="", head="final_conv"):
    """sd: state_dict-like mapping; x: (B,2,H,W). Returns (B,1,H,W). If training, BatchNorm uses batch statistics
    and `new_stats` (dict) receives the updated running statistics. `prefix` / `head` select a UNetStage inside a
    ProgressiveUNet state_dict (e.g. prefix="unet1.", head="final", ModelLoader.py:148-226)."""
    skips = []
    for name in ENCODERS:
        x = _block(sd, prefix + name, x, training, new_stats)
        skips.append(x)
        x = F.max_pool2d(x, kernel_size=2, stride=2)
    x = _block(sd, prefix + "bottleneck", x, training, new_stats)
    for (up, dec), skip in zip(DECODERS, reversed(skips)):
        x = F.conv_transpose2d(x, sd[f"{prefix}{up}.weight"], sd[f"{prefix}{up}.bias"], stride=2)
        x = torch.cat([x, skip], dim=1)
        x = _block(sd, prefix + dec, x, training, new_stats)
    return F.conv2d(x, sd[f"{prefix}{head}.weight"], sd[f"{prefix}{head}.bias"])


PROGRESSIVE_LOSS_WEIGHTS = (0.5, 1.0, 0.5)  # results/progressive_unet_history.json config.loss_weights


def progressive_forward(sd, slices, training=False, new_stats=None):
    """ProgressiveUNet.forward (ModelLoader.py:246-269): slices (B,5,H,W) -> (pred_i+1, pred_i+2, pred_i+3); the stage-1
    prediction feeds stages 2A/2B without detach."""
    i, i4 = slices[:, 0:1], slices[:, 4:5]
    p2 = unet_forward(sd, torch.cat([i, i4], dim=1), training, new_stats, "unet1.", "final")
    p1 = unet_forward(sd, torch.cat([i, p2], dim=1), training, new_stats, "unet2.", "final")
    p3 = unet_forward(sd, torch.cat([p2, i4], dim=1), training, new_stats, "unet3.", "final")
    return p1, p2, p3


def progressive_loss_and_grads(sd, slices, weights=PROGRESSIVE_LOSS_WEIGHTS):
    """Multi-scale MSE of the 3-stage chain and all parameter gradients (train mode)."""
    names = param_names(sd)
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in set(names) else v) for k, v in sd.items()}
    new_stats = {}
    p1, p2, p3 = progressive_forward(leaf, slices, True, new_stats)
    loss = sum(w * F.mse_loss(p, slices[:, k:k + 1]) for w, p, k in zip(weights, (p1, p2, p3), (1, 2, 3)))
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), (p1.detach(), p2.detach(), p3.detach()), dict(zip(names, grads)), new_stats


def param_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def loss_and_grads(sd, x, y, loss_fn=None):
    """Train-mode forward + backward. Returns (loss, output, {name: grad}, new running stats)."""
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in set(param_names(sd)) else v) for k, v in sd.items()}
    new_stats = {}
    out = unet_forward(leaf, x, training=True, new_stats=new_stats)
    loss = F.mse_loss(out, y) if loss_fn is None else loss_fn(out, y)
    names = param_names(sd)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), out.detach(), dict(zip(names, grads)), new_stats


def _bn(sd, name, x, training, new_stats):
    rm, rv = sd[f"{name}.running_mean"], sd[f"{name}.running_var"]
    if training:
        rm_new, rv_new = rm.detach().clone(), rv.detach().clone()
        x = F.batch_norm(x, rm_new, rv_new, sd[f"{name}.weight"], sd[f"{name}.bias"], training=True,
                         momentum=BN_MOMENTUM, eps=BN_EPS)
        if new_stats is not None:
            new_stats[f"{name}.running_mean"], new_stats[f"{name}.running_var"] = rm_new, rv_new
        return x
    return F.batch_norm(x, rm, rv, sd[f"{name}.weight"], sd[f"{name}.bias"], training=False, momentum=BN_MOMENTUM,
                        eps=BN_EPS)


def deepcnn_forward(sd, x, training=False, new_stats=None, num_blocks=(2, 2, 2, 2)):
    """DeepCNN.forward (ModelLoader.py:361-377) with ResidualBlock.forward (:290-307): 7x7 stem + BN + ReLU +
    MaxPool(3,1,1), four layers of residual blocks (all stride 1; 1x1 conv + BN downsample branch when the channel
    count changes), 1x1 output conv. The declared avgpool is not used by the reference forward."""
    x = F.conv2d(x, sd["conv1.weight"], None, padding=3)
    x = torch.relu(_bn(sd, "bn1", x, training, new_stats))
    x = F.max_pool2d(x, kernel_size=3, stride=1, padding=1)
    for li, nb in enumerate(num_blocks, start=1):
        for bi in range(nb):
            pre = f"layer{li}.{bi}"
            idn = x
            out = F.conv2d(x, sd[f"{pre}.conv1.weight"], None, padding=1)
            out = torch.relu(_bn(sd, f"{pre}.bn1", out, training, new_stats))
            out = F.conv2d(out, sd[f"{pre}.conv2.weight"], None, padding=1)
            out = _bn(sd, f"{pre}.bn2", out, training, new_stats)
            if f"{pre}.downsample.0.weight" in sd:
                idn = _bn(sd, f"{pre}.downsample.1", F.conv2d(x, sd[f"{pre}.downsample.0.weight"]), training, new_stats)
            x = torch.relu(out + idn)
    return F.conv2d(x, sd["output_conv.weight"], sd["output_conv.bias"])


def deepcnn_loss_and_grads(sd, x, y):
    names = param_names(sd)
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in set(names) else v) for k, v in sd.items()}
    new_stats = {}
    out = deepcnn_forward(leaf, x, True, new_stats)
    loss = F.mse_loss(out, y)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), out.detach(), dict(zip(names, grads)), new_stats


def adam_update(p, g, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """One torch.optim.Adam step (defaults of unet_model.py:155) on plain tensors; returns (p, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    return p - (lr / bc1) * m / denom, m, v


def train_step_cpu(sd, x, y, opt_state, step, lr=1e-4, loss_fn=None):
    """Reference train step on a state_dict (used as the timed CPU baseline in bench.py)."""
    loss, _, grads, new_stats = loss_and_grads(sd, x, y, loss_fn)
    with torch.no_grad():
        for k, g in grads.items():
            m, v = opt_state.setdefault(k, (torch.zeros_like(g), torch.zeros_like(g)))
            p, m, v = adam_update(sd[k], g, m, v, step, lr)
            sd[k] = p
            opt_state[k] = (m, v)
        sd.update(new_stats)
    return loss


def layer_taps(sd, x, y, loss_fn=None):
    """Train-mode forward + backward that also returns, for every layer, the tensors a per-layer (teacher-forced) parity
    test needs: the layer input, the raw conv output z (bias included, as nn.Conv2d produces it), the post-ReLU
    activation, and the gradients of the loss w.r.t. z and the activation. Used by tests/test_gpu_layers.py.
    Returns (loss, out, taps, grads) with taps[name] = dict(a_in, z, act, dz, dact) for the 18 Conv+BN+ReLU layers
    ('enc1.conv.0' ...) and taps['upconvK'] = dict(a_in, out, dout) for the 4 ConvTranspose2d layers."""
    names = param_names(sd)
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in set(names) else v) for k, v in sd.items()}
    taps = {}

    def block(prefix, t):
        for conv_i, bn_i in ((0, 1), (3, 4)):
            a_in = t
            z = F.conv2d(t, leaf[f"{prefix}.conv.{conv_i}.weight"], leaf.get(f"{prefix}.conv.{conv_i}.bias"), padding=1)
            rm = sd[f"{prefix}.conv.{bn_i}.running_mean"].detach().clone()
            rv = sd[f"{prefix}.conv.{bn_i}.running_var"].detach().clone()
            t = torch.relu(F.batch_norm(z, rm, rv, leaf[f"{prefix}.conv.{bn_i}.weight"], leaf[f"{prefix}.conv.{bn_i}.bias"],
                                        training=True, momentum=BN_MOMENTUM, eps=BN_EPS))
            taps[f"{prefix}.conv.{conv_i}"] = {"a_in": a_in, "z": z, "act": t}
        return t

    t = x
    skips = []
    for name in ENCODERS:
        t = block(name, t)
        skips.append(t)
        t = F.max_pool2d(t, kernel_size=2, stride=2)
    t = block("bottleneck", t)
    for (up, dec), skip in zip(DECODERS, reversed(skips)):
        u = F.conv_transpose2d(t, leaf[f"{up}.weight"], leaf[f"{up}.bias"], stride=2)
        taps[up] = {"a_in": t, "out": u}
        t = block(dec, torch.cat([u, skip], dim=1))
    out = F.conv2d(t, leaf["final_conv.weight"], leaf["final_conv.bias"])
    loss = F.mse_loss(out, y) if loss_fn is None else loss_fn(out, y)
    wanted = []
    for k, d in taps.items():
        wanted += [d["z"], d["act"]] if "z" in d else [d["out"]]
    gl = torch.autograd.grad(loss, wanted + [leaf[k] for k in names])
    it = iter(gl)
    for k, d in taps.items():
        if "z" in d:
            d["dz"], d["dact"] = next(it), next(it)
        else:
            d["dout"] = next(it)
    grads = dict(zip(names, it))
    for d in taps.values():
        for k in list(d):
            d[k] = d[k].detach()
    return loss.detach(), out.detach(), taps, grads
