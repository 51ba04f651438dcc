"""MSE + windowed-SSIM training loss on the fused CUDA kernel (b200sr_mse_ssim).

Reference: nn.MSELoss in UNetTrainer (unet_model.py:156,180). The combined-loss notebook
(notebooks/UNet_Training.ipynb) is missing from the reference snapshot; the definition frozen in SURVEY.md §8(a11)
is used: total = mse_weight*MSE + ssim_weight*(1 - mean SSIM), with
  mode 'gaussian': 11x11 Gaussian window sigma 1.5, valid map, biased covariance
  mode 'uniform' : 7x7 uniform window, valid map, sample covariance (skimage defaults,
                   VolumeVisualization.py:256)
C1 = (0.01*L)^2, C2 = (0.03*L)^2, L = data_range.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr


def ssim_window(mode: str):
    """Separable 1-D window taps and covariance normalisation for an SSIM mode."""
    if mode in ("gaussian", "G"):
        k, sigma = 11, 1.5
        g = [math.exp(-((i - k // 2) ** 2) / (2.0 * sigma * sigma)) for i in range(k)]
        s = sum(g)
        return [v / s for v in g], 1.0
    if mode in ("uniform", "U"):
        k = 7
        npix = k * k
        return [1.0 / k] * k, npix / (npix - 1.0)
    raise ValueError(f"Unknown SSIM mode: {mode}. Choose from: ['gaussian', 'uniform']")


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, module):
        loss, grad = module.value_and_grad(pred, target, need_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None


class CombinedLoss(nn.Module):
    def __init__(self, mse_weight=1.0, ssim_weight=0.005, mode="gaussian", data_range=1.0):
        super().__init__()
        self.mse_weight, self.ssim_weight, self.mode, self.data_range = mse_weight, ssim_weight, mode, data_range
        self.win, self.cov_norm = ssim_window(mode)
        self.C1, self.C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
        self._win_c = (ctypes.c_float * len(self.win))(*self.win)
        self.last_components = None  # (mse, ssim) device scalars of the last call
        self._ws = {}  # device -> (partial sums, ticket counter) of the deterministic in-kernel reduction

    def value_and_grad(self, pred, target, need_grad=True):
        """Returns (loss 0-d fp32 device tensor, dloss/dpred or None). No host synchronisation."""
        if not pred.is_cuda:
            raise _lib.B200SRError("b200sr CombinedLoss runs on CUDA only; there is no CPU path")
        if pred.shape != target.shape or pred.dim() != 4 or pred.shape[1] != 1:
            raise _lib.B200SRError(f"expected pred/target (B,1,H,W), got {tuple(pred.shape)} / {tuple(target.shape)}")
        pred = pred.detach().contiguous().float()
        target = target.detach().contiguous().float()
        B, _, H, W = pred.shape
        K = len(self.win)
        # deterministic finish inside the kernel: per-CTA partial sums in a persistent workspace, the last CTA adds them in
        # CTA order and writes {loss, mse, mean SSIM}; no atomics, no host-side arithmetic, no extra launches
        nblocks = B * ((H + 31) // 32) * ((W + 31) // 32)
        ws = self._ws.get(pred.device)
        if ws is None or ws[0].numel() < 2 * nblocks:
            ws = (torch.empty(2 * nblocks, dtype=torch.float64, device=pred.device),
                  torch.zeros(1, dtype=torch.int32, device=pred.device))
            self._ws[pred.device] = ws
        out3 = torch.empty(4, dtype=torch.float32, device=pred.device)  # fresh per call: the caller may keep the loss
        grad = torch.empty_like(pred) if need_grad else None
        call("b200sr_mse_ssim_det", ptr(pred), ptr(target), ptr(grad), ptr(out3), B, H, W,
             ctypes.cast(self._win_c, ctypes.c_void_p), K, self.cov_norm, self.C1, self.C2, self.mse_weight,
             self.ssim_weight, ptr(ws[0]), ws[0].numel(), ptr(ws[1]), _lib.current_stream_ptr())
        self.last_components = (out3[1], out3[2])
        return out3[0], grad

    def forward(self, pred, target):
        if pred.requires_grad and torch.is_grad_enabled():
            return _LossFn.apply(pred, target, self)
        return self.value_and_grad(pred, target, need_grad=False)[0]
