"""Drop-in mirror of the reference `src/unet_model.py` on top of the b200sr CUDA library.

Same names, constructor signatures, parameter names/shapes and state_dict layout as the reference
(/root/reference/src/unet_model.py:22-118 `UNetBlock`, `UNet`; :121-145 `MRIDataset`; :148-298 `UNetTrainer`;
:301-310 `create_dummy_dataset`), so notebooks and checkpoints keep working. What changes is what runs when
`model(x)` is called on a CUDA tensor: the whole forward/backward is executed by hand-written sm_100a kernels
through the C ABI in include/b200sr.h. There is no torch/cuDNN fallback on that path.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import Dataset

from . import _lib


class UNetBlock(nn.Module):
    """Double convolution block with batch norm — parameter container (reference unet_model.py:22-36).

    The nn.Sequential layout (conv.0 Conv2d, conv.1 BatchNorm2d, conv.2 ReLU, conv.3 Conv2d, conv.4 BatchNorm2d,
    conv.5 ReLU) is kept so that state_dict keys match `conv.{0,3}.{weight,bias}` and
    `conv.{1,4}.{weight,bias,running_mean,running_var,num_batches_tracked}`. The modules hold the parameters;
    the arithmetic is done by the engine.
    """

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        raise _lib.B200SRError(
            "UNetBlock is a parameter container in b200sr: call the parent UNet, whose forward runs the fused "
            "sm_100a kernels (a stand-alone block forward would be a torch fallback, which this package does not ship)")


class _UNetFunction(torch.autograd.Function):
    """Whole-network autograd node: forward and backward are each one pass through the CUDA engine."""

    @staticmethod
    def forward(ctx, model, x, *params):
        ctx.model = model
        return model._get_engine().forward_train(x)

    @staticmethod
    def backward(ctx, dout):
        engine = ctx.model._get_engine()
        engine.backward(dout, want_dx=ctx.needs_input_grad[1])
        # hand autograd an independent copy: the flat gradient buffer is reused by the next step
        flat = engine.flat_g.clone()
        grads = [flat[off:off + p.numel()].view(p.shape) for p, off in zip(engine._params(), engine.p_off)]
        return (None, engine.dx_input, *grads)


class UNet(nn.Module):
    """UNet for slice interpolation: (B,2,H,W) prior+posterior slices -> (B,1,H,W) middle slice.

    Reference: unet_model.py:39-118. Module tree, registration order and default init are identical, hence
    identical `state_dict()` (136 entries, 31,042,945 parameters).
    """

    def __init__(self, in_channels=2, out_channels=1, init_features=64):
        super().__init__()
        features = init_features
        self.in_channels, self.out_channels, self.init_features = in_channels, out_channels, init_features
        # arithmetic of the eval-mode forward: 'bf16' (bf16 operands, fp32 accumulation; rel-L2 ~3e-3 vs the fp32 reference)
        # or 'fp32' (bf16x3 operand splitting on the same tensor-core kernels; rel-L2 <= 1e-4, ~3x the tensor work).
        # Not a parameter / buffer: state_dict stays the reference's. Training always runs the bf16 path.
        self.eval_precision = os.environ.get("B200SR_EVAL_PRECISION", "bf16")

        self.enc1 = UNetBlock(in_channels, features)
        self.pool1 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.enc2 = UNetBlock(features, features * 2)
        self.pool2 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.enc3 = UNetBlock(features * 2, features * 4)
        self.pool3 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.enc4 = UNetBlock(features * 4, features * 8)
        self.pool4 = nn.MaxPool2d(kernel_size=2, stride=2)

        self.bottleneck = UNetBlock(features * 8, features * 16)

        self.upconv4 = nn.ConvTranspose2d(features * 16, features * 8, kernel_size=2, stride=2)
        self.dec4 = UNetBlock(features * 16, features * 8)
        self.upconv3 = nn.ConvTranspose2d(features * 8, features * 4, kernel_size=2, stride=2)
        self.dec3 = UNetBlock(features * 8, features * 4)
        self.upconv2 = nn.ConvTranspose2d(features * 4, features * 2, kernel_size=2, stride=2)
        self.dec2 = UNetBlock(features * 4, features * 2)
        self.upconv1 = nn.ConvTranspose2d(features * 2, features, kernel_size=2, stride=2)
        self.dec1 = UNetBlock(features * 2, features)

        self.final_conv = nn.Conv2d(features, out_channels, kernel_size=1)

    # the engine owns device buffers; it is rebuilt on demand and never pickled / deep-copied with the module
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

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_engine":
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        # parameters of the copy must not alias the flat buffer of the original
        for p in new.parameters():
            p.data = p.data.clone()
        return new

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.B200SRError(
                "b200sr.UNet runs on CUDA sm_100a only (input is on CPU). There is deliberately no CPU/torch "
                "fallback; use the reference implementation for CPU execution.")
        engine = self._get_engine()
        if self.training:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                return _UNetFunction.apply(self, x, *self.parameters())
            return engine.forward_train(x)
        return engine.forward_eval(x, precision=self.eval_precision)

    def set_eval_precision(self, precision):
        """'bf16' or 'fp32' (see __init__); returns self so it chains like .eval()."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"Unknown precision: {precision}. Choose from: ['bf16', 'fp32']")
        self.eval_precision = precision
        return self


class MRIDataset(Dataset):
    """Dataset wrapper for MRI triplets (reference unet_model.py:121-145): (prior, middle, posterior) ->
    input stack([prior, posterior]) (2,H,W), target middle (1,H,W)."""

    def __init__(self, triplets):
        self.triplets = triplets

    def __len__(self):
        return len(self.triplets)

    def __getitem__(self, idx):
        prior, middle, posterior = self.triplets[idx]
        input_data = np.stack([prior, posterior], axis=0).astype(np.float32)
        target_data = np.expand_dims(middle, axis=0).astype(np.float32)
        return torch.from_numpy(input_data), torch.from_numpy(target_data)


def create_dummy_dataset(num_samples=100, img_size=256, seed=None):
    """iid N(0,1) triplets (reference unet_model.py:301-310); `seed` is an addition for reproducibility."""
    rng = np.random.default_rng(seed)
    return [tuple(rng.standard_normal((img_size, img_size)).astype(np.float32) for _ in range(3))
            for _ in range(num_samples)]


from .data import DevicePrefetcher, unpack_batch as _unpack_batch  # noqa: E402


class UNetTrainer:
    """Trainer with the reference's interface (unet_model.py:148-298).

    `loss='mse'` reproduces the reference step (MSE + Adam lr 1e-4); `loss='combined'` trains with
    MSE + ssim_weight*(1-SSIM). The step itself (`train_step`) bypasses autograd: forward,
    fused loss+gradient kernel, hand-written backward into the flat gradient buffer, bucketed NCCL all-reduce
    overlapped with the rest of backward when torch.distributed is initialised, then Adam.
    """

    def __init__(self, model, device='cuda' if torch.cuda.is_available() else 'cpu', learning_rate=1e-4,
                 model_save_dir='models', loss='mse', ssim_weight=0.005, ssim_mode='gaussian', data_range=1.0,
                 verbose=True, use_cuda_graph=False, perceptual_weight=0.01, vgg_weights=None):
        from .losses import CombinedLoss
        self.model = model.to(device)
        self.device = device
        self.perceptual = None
        from .optim import FlatAdam
        self.optimizer = FlatAdam(self.model, lr=learning_rate)
        if loss == 'mse':
            self.criterion = CombinedLoss(mse_weight=1.0, ssim_weight=0.0, mode=ssim_mode, data_range=data_range)
        elif loss == 'combined':
            self.criterion = CombinedLoss(mse_weight=1.0, ssim_weight=ssim_weight, mode=ssim_mode,
                                          data_range=data_range)
        elif loss == 'combined_perceptual':
            # the full combined loss of README.md:82-86: MSE + perceptual (VGG16) + SSIM
            from .perceptual import PerceptualLoss, VGG16Features
            self.criterion = CombinedLoss(mse_weight=1.0, ssim_weight=ssim_weight, mode=ssim_mode,
                                          data_range=data_range)
            vgg = VGG16Features()
            if vgg_weights is not None:
                vgg.load_pretrained(vgg_weights)
            self.perceptual = PerceptualLoss(weight=perceptual_weight, vgg=vgg)
        else:
            raise ValueError(f"Unknown loss: {loss}. Choose from: ['mse', 'combined', 'combined_perceptual']")
        self.model_save_dir = Path(model_save_dir)
        self.model_save_dir.mkdir(parents=True, exist_ok=True)
        self.train_losses = []
        self.val_losses = []
        self.best_val_loss = float('inf')
        self.patience_counter = 0
        self.verbose = verbose
        self._reducer = None
        # optional: replay single-GPU steps from a CUDA graph (captured on the third call per input shape), so the ~130
        # kernel launches of a step cost one graph launch on the host. Off by default: the step is GPU-bound and the
        # host stays ahead of the device, measured 13.3 ms (graph) vs 13.1 ms (eager) per B=32 step
        import os
        self.use_cuda_graph = use_cuda_graph and os.environ.get("B200SR_NO_GRAPH") is None
        self._graphs, self._graph_calls = {}, {}
        if torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            from .ddp import broadcast_module_state
            broadcast_module_state(self.model)
        if verbose:
            print(f"Model initialized on device: {device}")
            print(f"Total parameters: {sum(p.numel() for p in self.model.parameters()):,}")

    # -- one optimisation step on device tensors; returns the loss as a 0-d device tensor (no host sync) --
    def _device_step(self, inputs, targets):
        """Everything of a step that runs on the device: forward, fused loss+gradient, backward (+ bucketed
        all-reduce when distributed), Adam kernel. No host-side state that changes from step to step."""
        engine = self.model._get_engine()
        out = engine.forward_train(inputs)
        loss, dout = self.criterion.value_and_grad(out, targets)
        if self.perceptual is not None:
            lp, gp = self.perceptual.value_and_grad(out, targets)
            loss = loss + lp
            dout = dout + gp
        hook, scale = None, 1.0
        from .ddp import is_distributed
        if is_distributed():
            if self._reducer is None or self._reducer.flat is not engine.flat_g:
                from .ddp import BucketReducer
                self._reducer = BucketReducer(engine.flat_g)
            hook = self._reducer.reduce_range
            scale = 1.0 / self._reducer.world_size
        engine.backward(dout, bucket_hook=hook)
        if hook is not None:
            self._reducer.wait()
        self.optimizer.device_step(grad_scale=scale)
        return loss

    def train_step(self, inputs, targets):
        self.model.train()
        from .ddp import is_distributed
        if self.use_cuda_graph and inputs.is_cuda and not is_distributed():
            key = (tuple(inputs.shape), tuple(targets.shape), inputs.device)
            g = self._graphs.get(key)
            if g is None:
                n = self._graph_calls.get(key, 0)
                self._graph_calls[key] = n + 1
                if n >= 2:   # two eager steps first: lazy one-time set-up (buffers, kernel attributes) must be done
                    g = self._capture(inputs, targets)
                    self._graphs[key] = g
            if g is not None:
                g["x"].copy_(inputs)
                g["y"].copy_(targets)
                self.optimizer.host_pre_step()
                g["graph"].replay()
                _lib.LAUNCH_COUNTER["n"] += g["launches"]
                return g["loss"]
        self.optimizer.host_pre_step()
        return self._device_step(inputs.contiguous().float(), targets.contiguous().float())

    def _capture(self, inputs, targets):
        x = inputs.detach().contiguous().float().clone()
        y = targets.detach().contiguous().float().clone()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.LAUNCH_COUNTER["n"]
        with torch.cuda.graph(graph):
            loss = self._device_step(x, y)
        launches = _lib.LAUNCH_COUNTER["n"] - n0
        _lib.LAUNCH_COUNTER["n"] = n0
        return {"graph": graph, "x": x, "y": y, "loss": loss, "launches": launches}

    def train_epoch(self, train_loader):
        self.model.train()
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        n = 0
        for inputs, targets in DevicePrefetcher(train_loader, self.device):  # H2D of batch i+1 overlaps step i
            total += self.train_step(inputs, targets).detach()
            n += 1
        return float(total.item()) / max(n, 1)  # single host sync per epoch (reference: one per step, :187)

    def validate(self, val_loader):
        self.model.eval()
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        n = 0
        with torch.no_grad():
            for batch in val_loader:
                inputs, targets = _unpack_batch(batch)
                inputs = inputs.to(self.device, non_blocking=True)
                targets = targets.to(self.device, non_blocking=True)
                outputs = self.model(inputs)
                loss, _ = self.criterion.value_and_grad(outputs, targets, need_grad=False)
                total += loss
                n += 1
        return float(total.item()) / max(n, 1)

    def train(self, train_loader, val_loader, epochs=100, early_stopping_patience=15):
        for epoch in range(1, epochs + 1):
            train_loss = self.train_epoch(train_loader)
            val_loss = self.validate(val_loader)
            self.train_losses.append(train_loss)
            self.val_losses.append(val_loss)
            msg = f"Epoch {epoch}/{epochs} | Train Loss: {train_loss:.4f} | Val Loss: {val_loss:.4f}"
            if val_loss < self.best_val_loss:
                self.best_val_loss = val_loss
                self.patience_counter = 0
                self.save_checkpoint(epoch, val_loss, is_best=True)
                msg += " (Best)"
            else:
                self.patience_counter += 1
                msg += f" (patience: {self.patience_counter}/{early_stopping_patience})"
            if self.verbose:
                print(msg)
            if self.patience_counter >= early_stopping_patience:
                if self.verbose:
                    print(f"Early stopping triggered after {epoch} epochs")
                break
        self.save_training_logs()

    def save_checkpoint(self, epoch, val_loss, is_best=False):
        """Same dict layout as the reference (unet_model.py:249-256) so ModelLoader.load_model reads it."""
        checkpoint = {
            'epoch': epoch,
            'model_state_dict': self.model.state_dict(),
            'optimizer_state_dict': self.optimizer.state_dict(),
            'val_loss': val_loss,
            'train_losses': self.train_losses,
            'val_losses': self.val_losses,
        }
        if is_best:
            torch.save(checkpoint, self.model_save_dir / 'unet_best.pt')
        torch.save(checkpoint, self.model_save_dir / 'unet_latest.pt')

    def save_training_logs(self):
        history = {'train_losses': self.train_losses, 'val_losses': self.val_losses,
                   'best_val_loss': self.best_val_loss}
        with open(self.model_save_dir / 'training_history.json', 'w') as f:
            json.dump(history, f, indent=4)
