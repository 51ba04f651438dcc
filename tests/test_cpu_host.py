"""CPU tests of the host side: drop-in surface (names, signatures, state_dict layout), loader contract, C-ABI
library loading/exports, loud failure without a GPU, synthetic data, bucket planning."""
import inspect
import os
import re

import numpy as np
import pytest
import torch

import b200sr
from b200sr import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "b200sr.h")).read()
    declared = set(re.findall(r"\b(b200sr_[a-zA-Z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in include/b200sr.h"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200sr.h but not exported by libb200sr.so"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.b200sr_version() >= 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.b200sr_device_ok() == 3  # B200SR_ENODEV
    assert "CUDA" in _lib.last_error() or "device" in _lib.last_error()
    m = b200sr.UNet()
    with pytest.raises(b200sr.B200SRError):
        m(torch.zeros(1, 2, 128, 256))
    with pytest.raises(b200sr.B200SRError):
        m.enc1(torch.zeros(1, 2, 16, 16))
    with pytest.raises(b200sr.B200SRError):
        b200sr.CombinedLoss()(torch.zeros(1, 1, 32, 32), torch.zeros(1, 1, 32, 32))


def test_bad_arguments_are_rejected_with_message():
    lib = _lib.load()
    rc = lib.b200sr_conv3x3_fwd(None, 64, 0, 64, None, 64, 1, 8, 16, None, 64, 0, None, None, 0, None, 0, None)
    assert rc == 1 and "b200sr_conv3x3_fwd" in _lib.last_error() or "requirement failed" in _lib.last_error()


def test_unet_constructor_and_state_dict_layout():
    sig = inspect.signature(b200sr.UNet.__init__)
    assert list(sig.parameters)[1:] == ["in_channels", "out_channels", "init_features"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [2, 1, 64]
    m = b200sr.UNet()
    sd = m.state_dict()
    assert len(sd) == 136
    assert sum(p.numel() for p in m.parameters()) == 31_042_945
    assert sum(b.numel() for b in m.buffers()) == 11_794
    top = []
    for k in sd:
        t = k.split(".")[0]
        if t not in top:
            top.append(t)
    assert top == ["enc1", "enc2", "enc3", "enc4", "bottleneck", "upconv4", "dec4", "upconv3", "dec3", "upconv2",
                   "dec2", "upconv1", "dec1", "final_conv"]
    assert sd["enc1.conv.0.weight"].shape == (64, 2, 3, 3) and sd["enc1.conv.0.bias"].shape == (64,)
    assert sd["bottleneck.conv.3.weight"].shape == (1024, 1024, 3, 3)
    assert sd["upconv4.weight"].shape == (1024, 512, 2, 2) and sd["upconv1.bias"].shape == (64,)
    assert sd["final_conv.weight"].shape == (1, 64, 1, 1)
    assert sd["enc1.conv.1.num_batches_tracked"].dtype == torch.int64
    assert all(v.dtype == torch.float32 for k, v in sd.items() if not k.endswith("num_batches_tracked"))
    m.load_state_dict(sd, strict=True)


def test_load_model_contract(tmp_path, capsys):
    with pytest.raises(ValueError, match="Unknown model"):
        b200sr.load_model("resnet", device="cpu", root=str(tmp_path))
    with pytest.raises(FileNotFoundError, match="Checkpoint not found"):
        b200sr.load_model("unet", device="cpu", root=str(tmp_path))
    torch.manual_seed(1)
    src = b200sr.UNet()
    sd = src.state_dict()
    (tmp_path / "models").mkdir()
    (tmp_path / "notebooks").mkdir()
    # layout 1: trainer checkpoint dict in models/
    torch.save({"epoch": 1, "model_state_dict": sd, "optimizer_state_dict": {}, "val_loss": 0.1,
                "train_losses": [], "val_losses": []}, tmp_path / "models" / "unet_best.pt")
    # layout 2: bare state_dict, found through the notebooks/ fallback
    torch.save(sd, tmp_path / "notebooks" / "unet_combined_best.pt")
    for name in ("unet", "UNET", "unet_combined"):
        m = b200sr.load_model(name, device="cpu", root=str(tmp_path), verbose=False)
        assert isinstance(m, b200sr.UNet) and not m.training
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k])
    # layout 3: GAN-style key
    torch.save({"generator_state_dict": sd}, tmp_path / "models" / "unet_best.pt")
    m = b200sr.load_model("unet", device="cpu", root=str(tmp_path), verbose=False)
    assert torch.equal(m.state_dict()["final_conv.bias"], sd["final_conv.bias"])
    # a checkpoint of another architecture under a registry name fails loudly (strict load), no silent fallback
    torch.save(sd, tmp_path / "models" / "unet_gan_best.pt")
    with pytest.raises(RuntimeError):
        b200sr.load_model("unet_gan", device="cpu", root=str(tmp_path))


def test_trainer_surface():
    sig = inspect.signature(b200sr.UNetTrainer.__init__)
    names = list(sig.parameters)
    assert names[1:5] == ["model", "device", "learning_rate", "model_save_dir"]
    assert sig.parameters["learning_rate"].default == 1e-4 and sig.parameters["model_save_dir"].default == "models"
    for meth in ("train_epoch", "validate", "train", "save_checkpoint", "save_training_logs", "train_step"):
        assert callable(getattr(b200sr.UNetTrainer, meth))
    with pytest.raises(ValueError):
        b200sr.UNetTrainer(b200sr.UNet(), device="cpu", loss="huber", model_save_dir="/tmp/b200sr_t", verbose=False)


def test_dataset_helpers():
    trip = b200sr.create_dummy_dataset(3, img_size=32, seed=0)
    ds = b200sr.MRIDataset(trip)
    x, y = ds[1]
    assert x.shape == (2, 32, 32) and y.shape == (1, 32, 32) and x.dtype == torch.float32
    assert np.array_equal(x[0].numpy(), trip[1][0]) and np.array_equal(x[1].numpy(), trip[1][2])
    assert np.array_equal(y[0].numpy(), trip[1][1])


def test_synthetic_generator_cpu():
    g1 = b200sr.SyntheticTripletGenerator(3, 64, 64, device="cpu", seed=7)
    g2 = b200sr.SyntheticTripletGenerator(3, 64, 64, device="cpu", seed=7)
    g3 = b200sr.SyntheticTripletGenerator(3, 64, 64, device="cpu", seed=7, rank=1)
    x1, y1 = g1.next()
    x2, y2 = g2.next()
    x3, _ = g3.next()
    assert x1.shape == (3, 2, 64, 64) and y1.shape == (3, 1, 64, 64)
    assert torch.equal(x1, x2) and torch.equal(y1, y2) and not torch.equal(x1, x3)
    assert torch.allclose(x1.mean(dim=(-2, -1)), torch.zeros(3, 2), atol=1e-4)
    assert torch.allclose(x1.std(dim=(-2, -1)), torch.ones(3, 2), atol=1e-3)
    # neighbouring slices are correlated: the target is closer to the input mean than to noise
    mid = 0.5 * (x1[:, 0] + x1[:, 1])
    assert float(((mid - y1[:, 0]) ** 2).mean()) < 0.5


def test_ssim_window():
    w, cn = b200sr.ssim_window("gaussian")
    assert len(w) == 11 and abs(sum(w) - 1) < 1e-12 and cn == 1.0
    w, cn = b200sr.ssim_window("uniform")
    assert len(w) == 7 and abs(cn - 49 / 48) < 1e-12
    with pytest.raises(ValueError):
        b200sr.ssim_window("box")


def test_shard_batch_and_bucket_coalescing():
    from b200sr.ddp import BucketReducer, shard_batch
    assert [shard_batch(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_batch(2, 3, 4) == (2, 2)
    flat = torch.zeros(100)
    red = BucketReducer(flat, min_bucket_elems=30)
    red.world_size = 2          # plan only: _launch is intercepted
    sent = []
    red._launch = lambda lo, hi: sent.append((lo, hi))
    for lo, hi in ((90, 100), (80, 90), (40, 80), (35, 40), (0, 35)):
        red.reduce_range(lo, hi)
    red.flush()
    assert sent == [(40, 100), (0, 40)], sent
