"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's evaluation metrics.

Follows /root/reference/src/VolumeVisualization.py:237-269 `compute_metrics(original, predicted)` line by line:
min-max normalisation by the ORIGINAL volume's range (+1e-8), prediction clipped to [0,1], per-slice SSIM and PSNR with
data_range=1.0, MAE over the volume, mean / population-std over the slices. The two scikit-image calls it makes
(skimage.metrics.structural_similarity / peak_signal_noise_ratio, scikit-image is not vendored in the reference and not
installed here) are restated from their published algorithms: SSIM = oracle.ssim_oracle.ssim_skimage_restatement
(7x7 uniform window, sample covariance, border crop; pinned in tests/test_cpu_oracle.py), PSNR = 10*log10(R^2 / MSE).
"""
from __future__ import annotations

import numpy as np

from .ssim_oracle import ssim_skimage_restatement


def psnr_restatement(image_true, image_test, data_range=1.0):
    err = np.mean((np.asarray(image_true, dtype=np.float64) - np.asarray(image_test, dtype=np.float64)) ** 2)
    return 10.0 * np.log10((data_range ** 2) / err)


def compute_metrics(original, predicted):
    original = np.asarray(original)
    predicted = np.asarray(predicted)
    orig_min = original.min()
    orig_max = original.max()
    orig_range = orig_max - orig_min + 1e-8
    orig_norm = (original - orig_min) / orig_range
    pred_norm = (predicted - orig_min) / orig_range
    pred_norm = np.clip(pred_norm, 0, 1)
    ssim_scores, psnr_scores = [], []
    for i in range(len(original)):
        ssim_scores.append(ssim_skimage_restatement(orig_norm[i], pred_norm[i], data_range=1.0))
        psnr_scores.append(psnr_restatement(orig_norm[i], pred_norm[i], data_range=1.0))
    mae = np.mean(np.abs(orig_norm - pred_norm))
    return {"ssim_mean": np.mean(ssim_scores), "ssim_std": np.std(ssim_scores), "psnr_mean": np.mean(psnr_scores),
            "psnr_std": np.std(psnr_scores), "mae": mae, "orig_norm": orig_norm, "pred_norm": pred_norm,
            "ssim_scores": np.asarray(ssim_scores), "psnr_scores": np.asarray(psnr_scores)}
