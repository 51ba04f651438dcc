"""CPU tests for the Fast-DDPM row: drop-in surface, scheduler, oracle vs golden (from the unmodified
reference), and the algebra behind the engine's time-embedding fold (checked on the oracle, no GPU needed)."""
import inspect
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import b200sr
from oracle import cases, fastddpm_oracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastddpm_golden.npz"))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_constructors_and_state_dict_layout():
    sig = inspect.signature(b200sr.FastDDPM.__init__)
    assert list(sig.parameters)[1:] == ["T", "device"] and sig.parameters["T"].default == 10
    sig = inspect.signature(b200sr.UNet2D.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [("in_ch", 3), ("base_ch", 64), ("time_dim", 256)]
    m = b200sr.FastDDPM(T=10, device="cpu")
    sd = m.state_dict()
    assert list(sd) == list(GOLD["keys"]) and len(sd) == 26
    assert sum(p.numel() for p in m.parameters()) == 2_162_177
    assert sd["unet.inc.block.0.weight"].shape == (64, 259, 3, 3) and sd["unet.up1.block.0.weight"].shape == (64, 192, 3, 3)
    assert sd["unet.time_mlp.2.weight"].shape == (256, 256) and sd["unet.outc.weight"].shape == (1, 64, 1, 1)
    with pytest.raises(b200sr.B200SRError):  # no CPU path
        m(torch.zeros(1, 2, 64, 64), torch.zeros(1, 1, 64, 64), torch.zeros(1, dtype=torch.long))
    with pytest.raises(b200sr.B200SRError):
        m.unet(torch.zeros(1, 3, 64, 64), torch.zeros(1, dtype=torch.long))


def test_scheduler_matches_reference_steps():
    s = b200sr.FastNoiseScheduler(10, "cpu")
    # SURVEY Appendix B / FastDDPM notebook: 10 of 1000 steps, 40 % up to t=699, 60 % after
    assert s.idxs.tolist() == [0, 233, 466, 699, 699, 759, 819, 879, 939, 999]
    np.testing.assert_array_equal(s.alpha_bar.numpy(), GOLD["alpha_bar"])
    np.testing.assert_allclose(s.coef_table[:, 0].numpy() ** 2 + s.coef_table[:, 1].numpy() ** 2, 1.0, rtol=1e-6)
    t = torch.tensor([0, 5, 9])
    assert torch.equal(b200sr.sinusoidal_timestep_embedding(t, 256), fastddpm_oracle.timestep_embedding(t, 256))
    # other chain lengths: mirror == oracle restatement == reference (golden vector for T=20)
    for T in (4, 5, 20, 50):
        assert torch.equal(b200sr.FastNoiseScheduler(T, "cpu").alpha_bar, fastddpm_oracle.schedule(T)[0])
    np.testing.assert_array_equal(b200sr.FastNoiseScheduler(20, "cpu").alpha_bar.numpy(), GOLD["alpha_bar_T20"])


def test_oracle_matches_golden():
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    cond, target, t, noise = cases.fastddpm_inputs()
    loss, eps, grads = fastddpm_oracle.loss_and_grads(sd, cond, target, t, noise)
    assert abs(float(loss) - float(GOLD["loss"])) / float(GOLD["loss"]) < 1e-6
    assert rel(eps, torch.from_numpy(GOLD["eps"])) < 1e-5
    names = list(GOLD["grad_names"])
    norms = np.array([grads["unet." + k if not k.startswith("unet.") else k].double().norm().item() for k in names])
    np.testing.assert_allclose(norms, GOLD["grad_norms"], rtol=1e-3, atol=1e-9)
    for k in names:
        np.testing.assert_allclose(grads[k].reshape(-1)[:cases.GRAD_HEAD].numpy(), GOLD["grad_head/" + k], rtol=2e-3,
                                   atol=1e-7)
    c = cases.FASTDDPM_CASE
    torch.manual_seed(c["noise_seed"] + 1)
    x_T = torch.randn(c["B"], 1, c["H"], c["W"])
    s = fastddpm_oracle.sample(sd, cond, x_T)
    assert rel(s, torch.from_numpy(GOLD["sample"])) < 1e-5
    assert float(s.abs().max()) <= 1.0


def test_time_embedding_fold_is_exact_algebra():
    """The engine replaces the 256 tiled time channels of the first conv by a per-sample, per-border-class bias and
    gets their weight gradient from border sums of dz. Same sums, re-associated: check both on the CPU oracle."""
    torch.manual_seed(3)
    B, H, W = 2, 8, 12
    w = torch.randn(64, 259, 3, 3, dtype=torch.float64) * 0.05
    bias = torch.randn(64, dtype=torch.float64)
    img = torch.randn(B, 3, H, W, dtype=torch.float64)
    e = torch.randn(B, 256, dtype=torch.float64)
    full = F.conv2d(torch.cat([img, e[:, :, None, None].expand(B, 256, H, W)], 1), w, bias, padding=1)
    # forward fold
    P = torch.einsum("ockl,bc->bokl", w[:, 3:], e)  # per-tap contribution [B][64][3][3]
    folded = F.conv2d(img, w[:, :3], None, padding=1)
    for h in range(H):
        for x in range(W):
            khs = [k for k in range(3) if 0 <= h + k - 1 < H]
            kws = [k for k in range(3) if 0 <= x + k - 1 < W]
            folded[:, :, h, x] += bias + P[:, :, khs][:, :, :, kws].sum(dim=(2, 3))
    assert rel(folded, full) < 1e-12
    # backward fold: dW_time[co][c][tap] = sum_b e[b][c] * S[b][co][tap], S from total / border rows / columns / corners
    dz = torch.randn(B, 64, H, W, dtype=torch.float64)
    wl = w.clone().requires_grad_(True)
    out = F.conv2d(torch.cat([img, e[:, :, None, None].expand(B, 256, H, W)], 1), wl, bias, padding=1)
    (out * dz).sum().backward()
    T = dz.sum(dim=(2, 3))
    S = torch.zeros(B, 64, 3, 3, dtype=torch.float64)
    for kh in range(3):
        for kw in range(3):
            s = T.clone()
            er = 0 if kh == 0 else (H - 1 if kh == 2 else None)
            ec = 0 if kw == 0 else (W - 1 if kw == 2 else None)
            if er is not None:
                s -= dz[:, :, er, :].sum(-1)
            if ec is not None:
                s -= dz[:, :, :, ec].sum(-1)
            if er is not None and ec is not None:
                s += dz[:, :, er, ec]
            S[:, :, kh, kw] = s
    dW_time = torch.einsum("bc,bokl->ockl", e, S)
    assert rel(dW_time, wl.grad[:, 3:]) < 1e-12


def test_load_model_fastddpm_and_generator(tmp_path):
    torch.manual_seed(4)
    sd = b200sr.FastDDPM(T=10, device="cpu").state_dict()
    (tmp_path / "models").mkdir()
    torch.save(sd, tmp_path / "models" / "fastddpm_advanced_best.pth")
    m = b200sr.load_model("fastddpm", device="cpu", root=str(tmp_path), verbose=False)
    assert isinstance(m, b200sr.FastDDPM) and not m.training and m.scheduler.T == 10
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    g = b200sr.UNetGenerator()
    gsd = g.state_dict()
    assert len(gsd) == 118 and "final.weight" in gsd and "enc1.conv.0.bias" not in gsd
    torch.save({"generator_state_dict": gsd}, tmp_path / "models" / "unet_gan_best.pt")
    m = b200sr.load_model("unet_gan", device="cpu", root=str(tmp_path), verbose=False)
    assert isinstance(m, b200sr.UNetGenerator) and not m.training
    with pytest.raises(b200sr.B200SRError):
        m(torch.zeros(1, 2, 128, 256))
