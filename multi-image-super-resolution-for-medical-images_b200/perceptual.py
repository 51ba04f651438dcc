"""VGG16 perceptual term of the combined loss on the b200sr kernels — SURVEY.md §8(f) row 2, BASELINE configs[2].

Reference evidence: README.md:82-86 ("Loss = MSE + perceptual (VGG) + SSIM"), results/training_curves_combined.png
(weight ~0.01) and results/unet_gan_history.json (`lambda_perceptual`). The notebook that defined the term is missing
from the snapshot (/root/reference/.MISSING_LARGE_BLOBS:14), so the definition is FROZEN HERE (parity unpinned):

    L_perc(pred, target) = mean( (phi(pred) - phi(target))^2 ),   phi = torchvision VGG16 `features[:16]`
    (conv1_1 .. relu3_3: 7 Conv3x3+bias+ReLU, 2 MaxPool2d(2,2)), single-channel slices replicated to 3 channels,
    no ImageNet normalisation (inputs are z-scored MRI slices), VGG weights frozen.

Pretrained weights cannot be downloaded here; `VGG16Features.load_pretrained(path)` accepts a torchvision
`vgg16` state_dict, otherwise the module is initialised like torchvision's (kaiming-normal, fan_out) under a fixed seed.

Everything runs on kernels of the UNet hot path: conv3x3 forward with a bias+ReLU epilogue, conv3x3 dgrad, 2x2
max-pool forward/backward, the first-layer direct conv (the 3 replicated input channels fold into one by summing the
kernels), plus two small elementwise kernels (ReLU backward, feature MSE + gradient). pred and target go through the
stack as ONE batch of 2B images; only the pred half is back-propagated.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr
from .engine import PACK_CONV_DGRAD, PACK_CONV_FWD, _PACK_JOB_DTYPE, _align, _jobs_to_device

# torchvision.models.vgg16().features[:16]: (index in `features`, Cin, Cout) and pool positions
_CONVS = ((0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256))
_POOL_AFTER = {2, 7}  # max-pool follows conv index 2 (features[4]) and 7 (features[9])


class VGG16Features(nn.Module):
    """Parameter container with torchvision's key layout (`features.<idx>.{weight,bias}`), frozen."""

    def __init__(self, seed=1234):
        super().__init__()
        layers = []
        cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256]
        cin = 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers)
        g = torch.Generator().manual_seed(seed)
        for m in self.features:
            if isinstance(m, nn.Conv2d):  # torchvision init: kaiming_normal_(fan_out, relu), bias 0
                std = (2.0 / (m.out_channels * 9)) ** 0.5
                with torch.no_grad():
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
                    m.bias.zero_()
        for p in self.parameters():
            p.requires_grad_(False)

    def load_pretrained(self, path):
        sd = torch.load(path, map_location="cpu")
        own = self.state_dict()
        self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)
        return self

    def forward(self, x):
        raise _lib.B200SRError("VGG16Features is a parameter container in b200sr: use PerceptualLoss")


class PerceptualLoss(nn.Module):
    def __init__(self, weight=0.01, vgg=None):
        super().__init__()
        self.weight = float(weight)
        self.vgg = vgg if vgg is not None else VGG16Features()
        self._ready = None  # (device, flat packed weights ...)
        self._bufs = {}

    def _prepare(self, device):
        if self._ready == device:
            return
        _lib.require_device()
        self.vgg.to(device)
        convs = [self.vgg.features[i] for i, _, _ in _CONVS]
        # first layer: 3 replicated input channels == 1 channel with summed kernels; laid out as the (64,2,3,3)
        # parameter the direct first-layer kernels expect (second input channel unused: zero weights)
        w0 = convs[0].weight.detach().float()
        self.w1 = torch.zeros(64, 2, 3, 3, device=device)
        self.w1[:, 0] = w0.sum(dim=1)
        self.ones64 = torch.ones(64, device=device)
        self.biases = [c.bias.detach().float().contiguous() for c in convs]
        total, self.off_f, self.off_d = 0, {}, {}
        for k, (_, cin, cout) in enumerate(_CONVS[1:], start=1):
            n = cin * cout * 9
            self.off_f[k] = total
            total += _align(n)
            self.off_d[k] = total
            total += _align(n)
        self.wp = torch.zeros(total, dtype=torch.bfloat16, device=device)
        self.wsrc = [c.weight.detach().float().contiguous() for c in convs]
        jobs = np.zeros(2 * (len(_CONVS) - 1), dtype=_PACK_JOB_DTYPE)
        for j, (k, (_, cin, cout)) in enumerate(list(enumerate(_CONVS))[1:]):
            src = self.wsrc[k].data_ptr()
            jobs[2 * j] = (src, self.wp.data_ptr() + 2 * self.off_f[k], PACK_CONV_FWD, cout, cin, 0, cin * cout * 9)
            jobs[2 * j + 1] = (src, self.wp.data_ptr() + 2 * self.off_d[k], PACK_CONV_DGRAD, cout, cin, 0, cin * cout * 9)
        dj = _jobs_to_device(jobs, device)
        call("b200sr_pack_jobs", dj.data_ptr(), len(jobs), _lib.current_stream_ptr())
        torch.cuda.current_stream().synchronize()  # the job table is a temporary
        self._ready = device

    def _workspace(self, B, H, W, device):
        key = (B, H, W)
        b = self._bufs.get(key)
        if b is not None:
            return b
        if H % 16 != 0 or W % 16 != 0:
            raise _lib.B200SRError(f"PerceptualLoss needs H % 16 == 0 and W % 16 == 0 (got {H}x{W})")
        bf = torch.bfloat16
        N = 2 * B
        shapes = {1: (H, W, 64), 2: (H, W, 64), 3: (H // 2, W // 2, 128), 4: (H // 2, W // 2, 128),
                  5: (H // 4, W // 4, 256), 6: (H // 4, W // 4, 256), 7: (H // 4, W // 4, 256)}
        b = {"x2": torch.zeros(N, 2, H, W, dtype=torch.float32, device=device)}
        for k, (h, w, c) in shapes.items():
            b[f"a{k}"] = torch.empty(N, h, w, c, dtype=bf, device=device)
        b["p1"] = torch.empty(N, H // 2, W // 2, 64, dtype=bf, device=device)
        b["p2"] = torch.empty(N, H // 4, W // 4, 128, dtype=bf, device=device)
        big = B * H * W * 64
        b["g"] = [torch.empty(big, dtype=bf, device=device) for _ in range(2)]
        b["dx2"] = torch.empty(B, 2, H, W, dtype=torch.float32, device=device)
        self._bufs[key] = b
        return b

    def value_and_grad(self, pred, target, need_grad=True):
        """Returns (weight * L_perc as 0-d fp32 device tensor, d/dpred of it (B,1,H,W) fp32 or None)."""
        if not pred.is_cuda:
            raise _lib.B200SRError("b200sr PerceptualLoss runs on CUDA only; there is no CPU path")
        if pred.shape != target.shape or pred.dim() != 4 or pred.shape[1] != 1:
            raise _lib.B200SRError(f"expected pred/target (B,1,H,W), got {tuple(pred.shape)} / {tuple(target.shape)}")
        dev = pred.device
        self._prepare(dev)
        B, _, H, W = pred.shape
        N = 2 * B
        b = self._workspace(B, H, W, dev)
        st = _lib.current_stream_ptr()
        b["x2"][:B, 0] = pred.detach()[:, 0]
        b["x2"][B:, 0] = target.detach()[:, 0]
        a = {k: b[f"a{k}"] for k in range(1, 8)}

        def conv(k, src, dst, h, w):
            _, cin, cout = _CONVS[k]
            call("b200sr_conv3x3_fwd", ptr(src), cin, 0, cin, self.wp.data_ptr() + 2 * self.off_f[k], cout, N, h, w,
                 ptr(dst), cout, 0, None, ptr(self.biases[k]), 1, None, 0, st)

        call("b200sr_conv1_fwd", ptr(b["x2"]), ptr(self.w1), ptr(self.ones64), ptr(self.biases[0]), 1, ptr(a[1]), None, 0,
             N, H, W, st)
        conv(1, a[1], a[2], H, W)
        call("b200sr_maxpool2x2_fwd", ptr(a[2]), 64, 0, 64, ptr(b["p1"]), N, H, W, st)
        conv(2, b["p1"], a[3], H // 2, W // 2)
        conv(3, a[3], a[4], H // 2, W // 2)
        call("b200sr_maxpool2x2_fwd", ptr(a[4]), 128, 0, 128, ptr(b["p2"]), N, H // 2, W // 2, st)
        conv(4, b["p2"], a[5], H // 4, W // 4)
        conv(5, a[5], a[6], H // 4, W // 4)
        conv(6, a[6], a[7], H // 4, W // 4)

        nfeat = B * (H // 4) * (W // 4) * 256
        sums = torch.zeros(1, dtype=torch.float64, device=dev)
        g0, g1 = (t.data_ptr() for t in b["g"])
        half = a[7][B:]
        call("b200sr_feat_mse_grad", ptr(a[7]), ptr(half), g0 if need_grad else None, ptr(sums),
             self.weight * 2.0 / nfeat, nfeat, st)
        loss = (self.weight * sums[0] / float(nfeat)).float()
        if not need_grad:
            return loss, None

        def dgrad(k, dy, dx, h, w):
            _, cin, cout = _CONVS[k]
            call("b200sr_conv3x3_dgrad", dy, cout, 0, cout, self.wp.data_ptr() + 2 * self.off_d[k], cin, B, h, w, dx,
                 cin, 0, None, 0, st)

        def relu_bwd(dy, act, out, n):
            call("b200sr_relu_bwd", dy, ptr(act), out, n, st)

        def dgrad_relu(k, dy, dx, act, h, w):
            """dgrad of conv k with the ReLU mask of the layer below (its stored activation) applied in the epilogue"""
            _, cin, cout = _CONVS[k]
            call("b200sr_conv3x3_dgrad_relu", dy, cout, 0, cout, self.wp.data_ptr() + 2 * self.off_d[k], cin, B, h, w, dx,
                 cin, 0, ptr(act), cin, 0, None, 0, st)

        h4, w4, h2, w2 = H // 4, W // 4, H // 2, W // 2
        dgrad_relu(6, g0, g1, a[6], h4, w4)                       # -> d z6 (masked by a6 > 0 in the epilogue)
        dgrad_relu(5, g1, g0, a[5], h4, w4)
        dgrad(4, g0, g1, h4, w4)                                  # -> d p2
        call("b200sr_maxpool2x2_bwd", ptr(a[4]), 128, 0, g1, None, 0, 0, 128, g0, B, h2, w2, st)
        relu_bwd(g0, a[4], g0, B * h2 * w2 * 128)
        dgrad_relu(3, g0, g1, a[3], h2, w2)
        dgrad(2, g1, g0, h2, w2)                                  # -> d p1
        call("b200sr_maxpool2x2_bwd", ptr(a[2]), 64, 0, g0, None, 0, 0, 64, g1, B, H, W, st)
        relu_bwd(g1, a[2], g1, B * H * W * 64)
        dgrad_relu(1, g1, g0, a[1], H, W)
        call("b200sr_conv1_dgrad", g0, ptr(self.w1), ptr(b["dx2"]), B, H, W, st)
        return loss, b["dx2"][:, 0:1].contiguous()

    def forward(self, pred, target):
        return self.value_and_grad(pred, target, need_grad=False)[0]
