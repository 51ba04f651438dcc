"""TEST INFRASTRUCTURE ONLY — SSIM / combined-loss oracle.

The reference's combined-loss source (notebooks/UNet_Training.ipynb) is absent from the snapshot
(/root/reference/.MISSING_LARGE_BLOBS:14), so this part is PARITY UNPINNED against the reference; the definition
is the one frozen in SURVEY.md §8(a11):
  total = w_mse * MSE + w_ssim * (1 - mean(SSIM_map))
  mode 'gaussian': 11x11 Gaussian window (sigma 1.5) as a "valid" depthwise correlation, biased covariance
  mode 'uniform' : 7x7 uniform window, valid, sample covariance (x49/48): skimage.metrics.structural_similarity
                   defaults, the call the reference makes at src/VolumeVisualization.py:256
  C1 = (0.01 L)^2, C2 = (0.03 L)^2.
`ssim_skimage_restatement` re-derives mode 'uniform' with scipy.ndimage.uniform_filter + border crop exactly as
skimage's published algorithm does; tests pin the torch oracle against it.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def window_1d(mode):
    if mode in ("gaussian", "G"):
        k, sigma = 11, 1.5
        g = torch.tensor([math.exp(-((i - k // 2) ** 2) / (2 * sigma * sigma)) for i in range(k)], dtype=torch.float64)
        return g / g.sum(), 1.0
    if mode in ("uniform", "U"):
        k = 7
        return torch.full((k,), 1.0 / k, dtype=torch.float64), (k * k) / (k * k - 1.0)
    raise ValueError(mode)


def ssim_map(x, y, mode="gaussian", data_range=1.0):
    """x, y: (B,1,H,W). Returns the valid SSIM map (B,1,H-K+1,W-K+1) in the dtype of x."""
    w1, cov_norm = window_1d(mode)
    w1 = w1.to(device=x.device, dtype=x.dtype)
    k = w1.numel()
    w2 = (w1[:, None] * w1[None, :]).view(1, 1, k, k)
    filt = lambda t: F.conv2d(t, w2)
    mx, my = filt(x), filt(y)
    sxx = cov_norm * (filt(x * x) - mx * mx)
    syy = cov_norm * (filt(y * y) - my * my)
    sxy = cov_norm * (filt(x * y) - mx * my)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    return ((2 * mx * my + c1) * (2 * sxy + c2)) / ((mx * mx + my * my + c1) * (sxx + syy + c2))


def combined_loss(pred, target, w_mse=1.0, w_ssim=0.005, mode="gaussian", data_range=1.0):
    mse = F.mse_loss(pred, target)
    if w_ssim == 0:
        return w_mse * mse
    return w_mse * mse + w_ssim * (1.0 - ssim_map(pred, target, mode, data_range).mean())


def ssim_skimage_restatement(im1, im2, data_range=1.0):
    """Mean SSIM of two 2-D float arrays, skimage defaults (win 7, uniform filter, sample covariance, K1 .01,
    K2 .03, border crop (win-1)//2)."""
    from scipy.ndimage import uniform_filter
    im1 = np.asarray(im1, dtype=np.float64)
    im2 = np.asarray(im2, dtype=np.float64)
    win, ndim = 7, 2
    npix = win ** ndim
    cov_norm = npix / (npix - 1.0)
    ux, uy = uniform_filter(im1, size=win), uniform_filter(im2, size=win)
    uxx, uyy, uxy = uniform_filter(im1 * im1, size=win), uniform_filter(im2 * im2, size=win), \
        uniform_filter(im1 * im2, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())
