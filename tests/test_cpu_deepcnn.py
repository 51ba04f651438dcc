"""CPU tests for the DeepCNN row: drop-in surface and oracle vs golden (from the unmodified reference)."""
import inspect
import os

import numpy as np
import pytest
import torch

import b200sr
from oracle import cases, unet_oracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deepcnn_golden.npz"))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_constructor_and_state_dict_layout():
    sig = inspect.signature(b200sr.DeepCNN.__init__)
    assert list(sig.parameters)[1:] == ["in_channels", "out_channels", "num_blocks", "base_features"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [2, 1, [2, 2, 2, 2], 64]
    m = b200sr.DeepCNN()
    sd = m.state_dict()
    assert list(sd) == list(GOLD["keys"]) and len(sd) == 122
    assert sum(p.numel() for p in m.parameters()) == 11_173_889
    assert sd["conv1.weight"].shape == (64, 2, 7, 7) and sd["layer2.0.downsample.0.weight"].shape == (128, 64, 1, 1)
    assert sd["output_conv.weight"].shape == (1, 512, 1, 1) and "layer1.0.conv1.bias" not in sd


def test_load_model_deepcnn(tmp_path):
    torch.manual_seed(4)
    sd = b200sr.DeepCNN().state_dict()
    (tmp_path / "models").mkdir()
    torch.save({"model_state_dict": sd}, tmp_path / "models" / "deepcnn_best.pt")
    m = b200sr.load_model("deepcnn", device="cpu", root=str(tmp_path), verbose=False)
    assert isinstance(m, b200sr.DeepCNN) and not m.training
    with pytest.raises(b200sr.B200SRError):
        m(torch.zeros(1, 2, 32, 32))


def test_oracle_matches_golden():
    sd = cases.seeded_state_dict(b200sr.DeepCNN, seed=5)
    c = cases.DEEPCNN_CASE
    x, y = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    loss, out, grads, stats = unet_oracle.deepcnn_loss_and_grads(sd, x, y)
    assert abs(float(loss) - float(GOLD["loss"])) / float(GOLD["loss"]) < 1e-6
    assert rel(out, torch.from_numpy(GOLD["train_out"])) < 1e-5
    names = list(GOLD["grad_names"])
    norms = np.array([grads[k].double().norm().item() for k in names])
    np.testing.assert_allclose(norms, GOLD["grad_norms"], rtol=2e-3, atol=1e-9)
    sd = dict(sd)
    sd.update(stats)
    ev = unet_oracle.deepcnn_forward(sd, x, training=False)
    assert rel(ev, torch.from_numpy(GOLD["eval_out"])) < 1e-5
