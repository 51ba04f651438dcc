"""CPU tests: the oracle against the committed golden fixtures (generated from the unmodified reference by
oracle/make_golden.py) and the SSIM oracle against the skimage restatement."""
import os

import numpy as np
import pytest
import torch

import b200sr
from oracle import cases, ssim_oracle, unet_oracle

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "unet_golden.npz"))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.fixture(scope="module")
def train_result():
    torch.set_num_threads(os.cpu_count() or 1)
    sd = cases.seeded_state_dict(b200sr.UNet)
    c = cases.TRAIN_CASE
    x, y = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    return sd, x, y, unet_oracle.loss_and_grads(sd, x, y)


def test_seeded_weights_match_reference_checksums(train_result):
    sd = train_result[0]
    names = unet_oracle.param_names(sd)
    assert names == list(GOLDEN["grad_names"])
    sums = np.array([sd[k].double().sum().item() for k in names])
    np.testing.assert_allclose(sums, GOLDEN["param_sums"], rtol=1e-9, atol=1e-9)


def test_train_forward_loss_and_grads_match_golden(train_result):
    sd, x, y, (loss, out, grads, stats) = train_result
    assert rel(out, torch.from_numpy(GOLDEN["train_out"])) < 1e-5
    assert abs(float(loss) - float(GOLDEN["train_loss"])) < 1e-6
    names = list(GOLDEN["grad_names"])
    norms = np.array([grads[k].double().norm().item() for k in names])
    real = GOLDEN["grad_norms"] > 1e-6
    np.testing.assert_allclose(norms[real], GOLDEN["grad_norms"][real], rtol=2e-3)
    for i, k in enumerate(names):
        if real[i]:
            n = min(cases.GRAD_HEAD, grads[k].numel())
            got = grads[k].flatten()[:n].numpy()
            np.testing.assert_allclose(got, GOLDEN["grad_heads"][i][:n], rtol=5e-2,
                                       atol=1e-3 * float(np.abs(GOLDEN["grad_heads"][i][:n]).max() + 1e-12))
    stat_names = list(GOLDEN["stat_names"])
    got = np.concatenate([stats[k].numpy().ravel() for k in stat_names])
    np.testing.assert_allclose(got, GOLDEN["stats_after"], rtol=1e-4, atol=1e-6)


def test_adam_step_matches_golden(train_result):
    sd, _, _, (_, _, grads, _) = train_result
    names = list(GOLDEN["grad_names"])
    for i, k in enumerate(names):
        p, _, _ = unet_oracle.adam_update(sd[k], grads[k], torch.zeros_like(sd[k]), torch.zeros_like(sd[k]), 1)
        d = (p - sd[k]).double().norm().item()
        if GOLDEN["adam_delta_norms"][i] > 1e-7 and GOLDEN["grad_norms"][i] > 1e-6:
            assert abs(d - GOLDEN["adam_delta_norms"][i]) / GOLDEN["adam_delta_norms"][i] < 2e-2, k


def test_eval_forward_matches_golden(train_result):
    sd, _, _, (_, _, _, stats) = train_result
    sd = dict(sd)
    sd.update(stats)
    c = cases.EVAL_CASE
    xe, _ = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    out = unet_oracle.unet_forward(sd, xe, training=False)
    assert rel(out, torch.from_numpy(GOLDEN["eval_out"])) < 1e-5


@pytest.mark.parametrize("mode", ["gaussian", "uniform"])
def test_combined_loss_matches_golden(train_result, mode):
    sd, x, y, _ = train_result
    loss, _, grads, _ = unet_oracle.loss_and_grads(
        sd, x, y, lambda p, t: ssim_oracle.combined_loss(p, t, 1.0, 0.005, mode))
    assert abs(float(loss) - float(GOLDEN[f"combined_{mode}_loss"])) < 1e-6


def test_ssim_uniform_matches_skimage_restatement():
    c = cases.TRAIN_CASE
    x, y = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    a, b = x[0, 0].double(), (0.7 * x[0, 0] + 0.3 * y[0, 0]).double()
    s_t = float(ssim_oracle.ssim_map(a[None, None], b[None, None], "uniform").mean())
    s_s = ssim_oracle.ssim_skimage_restatement(a.numpy(), b.numpy())
    assert abs(s_t - s_s) < 1e-10
    assert abs(s_s - float(GOLDEN["ssim_uniform_pair"])) < 1e-12


def test_ssim_properties():
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, 1, 40, 48, generator=g, dtype=torch.float64)
    for mode, k in (("gaussian", 11), ("uniform", 7)):
        m = ssim_oracle.ssim_map(a, a, mode)
        assert m.shape == (2, 1, 40 - k + 1, 48 - k + 1)
        assert torch.allclose(m, torch.ones_like(m), atol=1e-12)  # SSIM(x,x) = 1
        b = torch.rand(2, 1, 40, 48, generator=g, dtype=torch.float64)
        assert torch.allclose(ssim_oracle.ssim_map(a, b, mode), ssim_oracle.ssim_map(b, a, mode))  # symmetric


def test_metrics_oracle_follows_the_reference_definition():
    """oracle/metrics_oracle.py restates compute_metrics (VolumeVisualization.py:237-269): normalisation by the ORIGINAL
    range, clipped prediction, per-slice SSIM (pinned restatement above) / PSNR, MAE, population std."""
    import numpy as np
    from oracle import metrics_oracle, ssim_oracle
    rng = np.random.default_rng(0)
    o = rng.standard_normal((4, 64, 48)).astype(np.float32)
    p = (o + 0.1 * rng.standard_normal(o.shape)).astype(np.float32)
    p[0] += 10.0   # far outside the original's range -> clipped to 1
    r = metrics_oracle.compute_metrics(o, p)
    lo, rng_ = o.min(), o.max() - o.min() + 1e-8
    on, pn = (o - lo) / rng_, np.clip((p - lo) / rng_, 0, 1)
    assert np.array_equal(r["orig_norm"], on) and np.array_equal(r["pred_norm"], pn)
    assert r["pred_norm"][0].max() == 1.0 and on.min() == 0.0
    assert abs(r["mae"] - np.abs(on - pn).mean()) < 1e-12
    s = [ssim_oracle.ssim_skimage_restatement(on[i], pn[i]) for i in range(4)]
    q = [10 * np.log10(1.0 / np.mean((on[i].astype(np.float64) - pn[i]) ** 2)) for i in range(4)]
    assert abs(r["ssim_mean"] - np.mean(s)) < 1e-12 and abs(r["ssim_std"] - np.std(s)) < 1e-12
    assert abs(r["psnr_mean"] - np.mean(q)) < 1e-9 and abs(r["psnr_std"] - np.std(q)) < 1e-9
    same = metrics_oracle.compute_metrics(o[1:], o[1:])
    assert abs(same["ssim_mean"] - 1.0) < 1e-12 and same["mae"] == 0.0
