"""Evaluation metrics of the reference on the device.

Mirror of `compute_metrics(original, predicted)` in /root/reference/src/VolumeVisualization.py:237-269 (same argument
meaning, same returned keys): min-max normalisation by the ORIGINAL volume's range (+1e-8), prediction clipped to [0,1],
per-slice SSIM with the scikit-image defaults the reference calls (7x7 uniform window, sample covariance, data_range 1) and
PSNR, MAE over the volume, mean / population std over the slices. One C-ABI call (b200sr_volume_metrics: min/max ->
normalise -> fused per-slice SSIM/PSNR/MAE kernel with a deterministic in-kernel finish); nothing is computed on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

_WS = {}


def compute_metrics(original, predicted, device="cuda"):
    """original, predicted: (S,H,W) volumes — numpy arrays (as in the reference) or torch tensors. Returns the reference's
    dict: 'ssim_mean', 'ssim_std', 'psnr_mean', 'psnr_std', 'mae' (Python floats), 'orig_norm', 'pred_norm' (numpy arrays
    for numpy inputs, device tensors for tensor inputs), plus 'ssim_scores' / 'psnr_scores' per slice."""
    as_numpy = isinstance(original, np.ndarray)
    o = torch.as_tensor(original)
    p = torch.as_tensor(predicted)
    if o.shape != p.shape or o.dim() != 3:
        raise _lib.B200SRError(f"expected two (S,H,W) volumes of equal shape, got {tuple(o.shape)} / {tuple(p.shape)}")
    if not o.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.B200SRError("b200sr.compute_metrics runs on CUDA sm_100a only; there is no CPU path")
        o, p = o.to(device), p.to(device)
    o, p = o.contiguous().float(), p.to(o.device).contiguous().float()
    S, H, W = o.shape
    nblocks = S * ((H + 31) // 32) * ((W + 31) // 32)
    ws = _WS.get(o.device)
    if ws is None or ws[0].numel() < 2048 + 4 * nblocks:
        ws = (torch.empty(2048 + 4 * nblocks, dtype=torch.float64, device=o.device),
              torch.zeros(2, dtype=torch.int32, device=o.device))
        _WS[o.device] = ws
    orig_norm, pred_norm = torch.empty_like(o), torch.empty_like(o)
    out5 = torch.empty(5, dtype=torch.float32, device=o.device)
    per_slice = torch.empty((S, 2), dtype=torch.float32, device=o.device)
    call("b200sr_volume_metrics", ptr(o), ptr(p), S, H, W, ptr(orig_norm), ptr(pred_norm), ptr(out5), ptr(per_slice),
         ptr(ws[0]), ws[0].numel(), ptr(ws[1]), _lib.current_stream_ptr())
    vals = out5.cpu()   # the one device->host read of the call
    res = {"ssim_mean": float(vals[0]), "ssim_std": float(vals[1]), "psnr_mean": float(vals[2]),
           "psnr_std": float(vals[3]), "mae": float(vals[4])}
    if as_numpy:
        res.update(orig_norm=orig_norm.cpu().numpy(), pred_norm=pred_norm.cpu().numpy(),
                   ssim_scores=per_slice[:, 0].cpu().numpy(), psnr_scores=per_slice[:, 1].cpu().numpy())
    else:
        res.update(orig_norm=orig_norm, pred_norm=pred_norm, ssim_scores=per_slice[:, 0], psnr_scores=per_slice[:, 1])
    return res
