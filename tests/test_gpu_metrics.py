"""On-device evaluation metrics (reference compute_metrics, src/VolumeVisualization.py:237-269) against the CPU oracle
(oracle/metrics_oracle.py: the reference function restated line by line, scikit-image's SSIM / PSNR from their published
algorithms; the SSIM restatement is pinned in tests/test_cpu_oracle.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _volume(S, H, W, seed):
    from oracle import cases
    sl = cases.seeded_slices((S + 4) // 5, H, W, seed).reshape(-1, H, W)[:S]
    g = torch.Generator().manual_seed(seed + 1)
    pred = sl + 0.15 * torch.randn(sl.shape, generator=g)
    pred[0] += 3.0    # drives part of the prediction out of the original's range: exercises the clip
    return sl.numpy().astype(np.float32), pred.numpy().astype(np.float32)


@pytest.mark.parametrize("shape", [(11, 256, 256), (5, 96, 80), (60, 256, 256)])
def test_compute_metrics_matches_oracle(shape):
    import b200sr
    from oracle import metrics_oracle
    orig, pred = _volume(*shape, seed=321)
    ref = metrics_oracle.compute_metrics(orig, pred)
    got = b200sr.compute_metrics(orig, pred)
    assert set(ref) <= set(got)
    for k in ("ssim_mean", "ssim_std", "psnr_mean", "psnr_std", "mae"):
        assert abs(got[k] - float(ref[k])) <= 1e-4 * max(abs(float(ref[k])), 1e-3), (k, got[k], float(ref[k]))
    assert isinstance(got["orig_norm"], np.ndarray)
    np.testing.assert_allclose(got["orig_norm"], ref["orig_norm"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(got["pred_norm"], ref["pred_norm"], rtol=0, atol=1e-6)
    assert got["pred_norm"].min() >= 0.0 and got["pred_norm"].max() <= 1.0
    np.testing.assert_allclose(got["ssim_scores"], ref["ssim_scores"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got["psnr_scores"], ref["psnr_scores"], rtol=1e-4)


def test_compute_metrics_is_reproducible_and_keeps_tensors_on_device():
    import b200sr
    orig, pred = _volume(7, 128, 256, seed=5)
    o, p = torch.from_numpy(orig).cuda(), torch.from_numpy(pred).cuda()
    a = b200sr.compute_metrics(o, p)
    b = b200sr.compute_metrics(o, p)
    assert a["orig_norm"].is_cuda and a["ssim_scores"].is_cuda
    for k in ("ssim_mean", "ssim_std", "psnr_mean", "psnr_std", "mae"):
        assert a[k] == b[k]
    assert torch.equal(a["ssim_scores"], b["ssim_scores"]) and torch.equal(a["psnr_scores"], b["psnr_scores"])


def test_compute_metrics_rejects_bad_input():
    import b200sr
    with pytest.raises(b200sr.B200SRError):
        b200sr.compute_metrics(np.zeros((2, 32, 32), np.float32), np.zeros((3, 32, 32), np.float32))
