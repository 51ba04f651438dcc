"""CPU tests for the Progressive UNet row: drop-in surface and oracle vs the golden fixtures generated
from the unmodified reference (oracle/make_golden.py)."""
import inspect
import os

import numpy as np
import pytest
import torch

import b200sr
from oracle import cases, unet_oracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "progressive_golden.npz"))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_state_dict_layout_matches_reference():
    m = b200sr.ProgressiveUNet(base_features=64)
    sd = m.state_dict()
    assert list(sd) == list(GOLD["keys"])
    assert len(sd) == 354 and sum(p.numel() for p in m.parameters()) == 93_111_171
    assert "unet1.enc1.conv.0.bias" not in sd and "unet2.final.weight" in sd      # bias=False convs, head named `final`
    assert list(inspect.signature(b200sr.UNetStage.__init__).parameters)[1:] == ["in_channels", "out_channels",
                                                                                 "base_features"]
    assert list(inspect.signature(b200sr.ProgressiveUNet.__init__).parameters)[1:] == ["base_features"]


def test_load_model_progressive(tmp_path):
    torch.manual_seed(2)
    sd = b200sr.ProgressiveUNet().state_dict()
    (tmp_path / "models").mkdir()
    torch.save({"model_state_dict": sd}, tmp_path / "models" / "progressive_unet_best.pt")
    m = b200sr.load_model("progressive_unet", device="cpu", root=str(tmp_path), verbose=False)
    assert isinstance(m, b200sr.ProgressiveUNet) and not m.training
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    with pytest.raises(b200sr.B200SRError):
        m(torch.zeros(1, 5, 128, 256))  # no CPU path


def test_oracle_matches_golden():
    torch.set_num_threads(os.cpu_count() or 1)
    sd = cases.seeded_state_dict(b200sr.ProgressiveUNet, seed=3)
    c = cases.PROGRESSIVE_CASE
    sl = cases.seeded_slices(c["B"], c["H"], c["W"], c["seed"])
    loss, (p1, p2, p3), grads, stats = unet_oracle.progressive_loss_and_grads(sd, sl)
    assert abs(float(loss) - float(GOLD["loss"])) < 1e-6
    for p, k in ((p1, "p1"), (p2, "p2"), (p3, "p3")):
        assert rel(p, torch.from_numpy(GOLD[k])) < 1e-5
    names = list(GOLD["grad_names"])
    norms = np.array([grads[k].double().norm().item() for k in names])
    np.testing.assert_allclose(norms, GOLD["grad_norms"], rtol=2e-3, atol=1e-9)
    sd = dict(sd)
    sd.update(stats)  # the golden eval pass ran after the train-mode forward had updated the running statistics
    e1, e2, e3 = unet_oracle.progressive_forward(sd, sl, training=False)
    for p, k in ((e1, "eval_p1"), (e2, "eval_p2"), (e3, "eval_p3")):
        assert rel(p, torch.from_numpy(GOLD[k])) < 1e-5
