"""Progressive UNet (3-stage 6 mm chain) on the b200sr kernels — SURVEY.md §8(f) row 1, BASELINE configs[3].

Mirror of the reference classes `ProgressiveUNetBlock`, `UNetStage`, `ProgressiveUNet`
(/root/reference/src/ModelLoader.py:33-47, :148-226, :229-269): same constructor signatures, module tree and
state_dict layout (`unet{1,2,3}.<stage keys>`, conv bias=False, 1x1 head named `final`; 354 entries, 93,111,171
parameters). Each stage runs on its own UNetEngine, i.e. on exactly the kernels of the UNet hot path; the only new
kernel is the first-layer data gradient (b200sr_conv1_dgrad), because stages 2A/2B receive stage 1's prediction as
an input channel and the reference does not detach it (ModelLoader.py:258-267).

Chain (reference forward): unet1(i, i+4) -> p2 ; unet2(i, p2) -> p1 ; unet3(p2, i+4) -> p3 ; returns (p1, p2, p3).
Loss (results/progressive_unet_history.json `config.loss_weights`): 0.5*MSE(p1) + 1.0*MSE(p2) + 0.5*MSE(p3).
"""
from __future__ import annotations

from pathlib import Path

import torch
import torch.nn as nn

from . import _lib
from .unet_model import _UNetFunction


class ProgressiveUNetBlock(nn.Module):
    """Parameter container (reference ModelLoader.py:33-47): (Conv3x3 bias=False -> BN -> ReLU) x 2."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        if kernel_size != 3 or stride != 1 or padding != 1:
            raise NotImplementedError("b200sr implements the reference block: kernel 3, stride 1, padding 1")
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        raise _lib.B200SRError("ProgressiveUNetBlock is a parameter container in b200sr: call the parent UNetStage")


class UNetStage(nn.Module):
    """One stage: (B,2,H,W) -> (B,1,H,W). Reference ModelLoader.py:148-226."""

    def __init__(self, in_channels=2, out_channels=1, base_features=64):
        super().__init__()
        f = base_features
        self.in_channels, self.out_channels, self.init_features = in_channels, out_channels, base_features
        self.enc1 = ProgressiveUNetBlock(in_channels, f)
        self.pool1 = nn.MaxPool2d(2, 2)
        self.enc2 = ProgressiveUNetBlock(f, f * 2)
        self.pool2 = nn.MaxPool2d(2, 2)
        self.enc3 = ProgressiveUNetBlock(f * 2, f * 4)
        self.pool3 = nn.MaxPool2d(2, 2)
        self.enc4 = ProgressiveUNetBlock(f * 4, f * 8)
        self.pool4 = nn.MaxPool2d(2, 2)
        self.bottleneck = ProgressiveUNetBlock(f * 8, f * 16)
        self.upconv4 = nn.ConvTranspose2d(f * 16, f * 8, kernel_size=2, stride=2)
        self.dec4 = ProgressiveUNetBlock(f * 16, f * 8)
        self.upconv3 = nn.ConvTranspose2d(f * 8, f * 4, kernel_size=2, stride=2)
        self.dec3 = ProgressiveUNetBlock(f * 8, f * 4)
        self.upconv2 = nn.ConvTranspose2d(f * 4, f * 2, kernel_size=2, stride=2)
        self.dec2 = ProgressiveUNetBlock(f * 4, f * 2)
        self.upconv1 = nn.ConvTranspose2d(f * 2, f, kernel_size=2, stride=2)
        self.dec1 = ProgressiveUNetBlock(f * 2, f)
        self.final = nn.Conv2d(f, out_channels, kernel_size=1)

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            from .engine import UNetEngine
            eng = UNetEngine(self)
            self.__dict__["_engine"] = eng
        return eng

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_engine", None)
        return state

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.B200SRError("b200sr.UNetStage runs on CUDA sm_100a only; there is no CPU/torch fallback")
        engine = self._get_engine()
        if self.training:
            if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
                return _UNetFunction.apply(self, x, *self.parameters())
            return engine.forward_train(x)
        return engine.forward_eval(x)


class GANUNetBlock(ProgressiveUNetBlock):
    """Reference ModelLoader.py:49-63: the same (Conv3x3 bias=False -> BN -> ReLU) x 2 block under its GAN name."""


class UNetGenerator(UNetStage):
    """Generator of the UNet-GAN (reference ModelLoader.py:383-463): structurally a UNetStage (bias-free blocks, head
    `final`, 118 state_dict entries), so it runs on the same UNetEngine. `load_model('unet_gan')` returns it for
    inference; the adversarial training loop is out of scope (the discriminator's source is missing from the reference
    snapshot), supervised fine-tuning works through UNetTrainer."""

    def __init__(self, in_channels=2, out_channels=1, base_features=64):
        super().__init__(in_channels, out_channels, base_features)


class ProgressiveUNet(nn.Module):
    """Reference ModelLoader.py:229-269. Input (B,5,H,W) slices i..i+4; returns (pred_i+1, pred_i+2, pred_i+3)."""

    def __init__(self, base_features=64):
        super().__init__()
        self.unet1 = UNetStage(2, 1, base_features)
        self.unet2 = UNetStage(2, 1, base_features)
        self.unet3 = UNetStage(2, 1, base_features)

    def forward(self, slices):
        i = slices[:, 0:1]
        i4 = slices[:, 4:5]
        p2 = self.unet1(torch.cat([i, i4], dim=1))
        p1 = self.unet2(torch.cat([i, p2], dim=1))
        p3 = self.unet3(torch.cat([p2, i4], dim=1))
        return p1, p2, p3


class ProgressiveUNetTrainer:
    """Train step of the 3-stage chain without autograd: three engine forwards, three fused MSE(+grad) kernels,
    backward through stages 2A/2B (which also yields the gradient w.r.t. stage 1's prediction), then stage 1 with
    the accumulated output gradient, then one flat Adam step per stage. Defaults follow
    results/progressive_unet_history.json (`lr 5e-4`, loss weights 0.5 / 1.0 / 0.5)."""

    def __init__(self, model, device="cuda", learning_rate=5e-4, loss_weights=(0.5, 1.0, 0.5), model_save_dir="models",
                 verbose=True):
        from .losses import CombinedLoss
        from .optim import FlatAdam
        self.model = model.to(device)
        self.device = device
        self.loss_weights = tuple(float(w) for w in loss_weights)
        self.criteria = [CombinedLoss(mse_weight=w, ssim_weight=0.0) for w in self.loss_weights]
        self.stages = [self.model.unet2, self.model.unet1, self.model.unet3]  # order of (p1, p2, p3)
        self.optimizers = [FlatAdam(s, lr=learning_rate) for s in (self.model.unet1, self.model.unet2, self.model.unet3)]
        self.model_save_dir = Path(model_save_dir)
        self.model_save_dir.mkdir(parents=True, exist_ok=True)
        self.last_losses = None
        if verbose:
            print(f"Total parameters: {sum(p.numel() for p in self.model.parameters()):,}")

    def train_step(self, slices):
        """slices: (B,5,H,W) fp32 on the device. Returns the weighted total loss (0-d device tensor, no host sync)."""
        self.model.train()
        m = self.model
        e1, e2, e3 = m.unet1._get_engine(), m.unet2._get_engine(), m.unet3._get_engine()
        i, t1, t2, t3, i4 = (slices[:, k:k + 1] for k in range(5))
        p2 = e1.forward_train(torch.cat([i, i4], dim=1))
        p1 = e2.forward_train(torch.cat([i, p2], dim=1))
        p3 = e3.forward_train(torch.cat([p2, i4], dim=1))
        l1, g1 = self.criteria[0].value_and_grad(p1, t1.contiguous())
        l2, g2 = self.criteria[1].value_and_grad(p2, t2.contiguous())
        l3, g3 = self.criteria[2].value_and_grad(p3, t3.contiguous())
        e2.backward(g1, want_dx=True)
        e3.backward(g3, want_dx=True)
        # stage 1's prediction is channel 1 of stage 2A's input and channel 0 of stage 2B's (no detach in the reference)
        g2_total = g2 + e2.dx_input[:, 1:2] + e3.dx_input[:, 0:1]
        e1.backward(g2_total)
        for opt in self.optimizers:
            opt.step()
        self.last_losses = (l1, l2, l3)
        return l1 + l2 + l3

    def save_checkpoint(self, epoch, val_loss, is_best=False):
        ck = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "val_loss": val_loss}
        if is_best:
            torch.save(ck, self.model_save_dir / "progressive_unet_best.pt")
        torch.save(ck, self.model_save_dir / "progressive_unet_latest.pt")
