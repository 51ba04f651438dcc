"""Adam over the engine's flat parameter / gradient buffers: one kernel launch for all 31 M parameters.

Reference: `optim.Adam(self.model.parameters(), lr=learning_rate)` in UNetTrainer.__init__ (unet_model.py:155) and
`self.optimizer.step()` (:185). Subclasses torch.optim.Adam so `state_dict()` / `load_state_dict()` keep the
reference checkpoint layout ('optimizer_state_dict', unet_model.py:252): the per-parameter `exp_avg` /
`exp_avg_sq` state tensors are views into two flat fp32 buffers that the kernel updates.

The step count lives ON THE DEVICE (`_step_dev`, int32[2]): the kernel derives the bias corrections 1-beta1^t and
sqrt(1-beta2^t) of the step it executes and increments the counter itself (b200sr_adam_step_auto). Nothing
step-dependent crosses from the host, so a host that runs several unsynchronised steps ahead of the device (the
normal state of the train loop) cannot hand step n the corrections of step n+k, and the launch replays from a CUDA
graph unchanged. The host only mirrors the count (`_nsteps`) for `state_dict()`.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr


class FlatAdam(torch.optim.Adam):
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(model.parameters(), lr=lr, betas=betas, eps=eps)
        self._model = model
        self._m = None
        self._v = None
        self._bound_to = None
        self._nsteps = 0
        self._step_dev = None

    def _bind(self, engine):
        """(Re)create the flat moment buffers for the engine's current flat parameter storage, carrying over whatever
        per-parameter state exists (a loaded checkpoint, or a previous binding)."""
        n = engine.p_total
        dev = engine.flat_p.device
        new_m = torch.zeros(n, dtype=torch.float32, device=dev)
        new_v = torch.zeros(n, dtype=torch.float32, device=dev)
        steps = 0
        for p, off in zip(engine._params(), engine.p_off):
            st = self.state.get(p)
            if st:  # carry over state loaded from a checkpoint or from a previous binding
                new_m[off:off + p.numel()].copy_(st["exp_avg"].reshape(-1))
                new_v[off:off + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                steps = max(steps, int(st["step"]))
            self.state[p] = {
                "step": torch.tensor(float(steps), dtype=torch.float32),
                "exp_avg": new_m[off:off + p.numel()].view(p.shape),
                "exp_avg_sq": new_v[off:off + p.numel()].view(p.shape),
            }
        if self._bound_to is not None and self._nsteps > steps:
            steps = self._nsteps  # re-binding after parameters moved: the per-parameter mirrors are refreshed lazily
        self._m, self._v = new_m, new_v
        self._nsteps = steps
        self._step_dev = torch.tensor([steps, 0], dtype=torch.int32, device=dev)
        self._bound_to = engine.flat_p.data_ptr()

    def _check_group(self):
        group = self.param_groups[0]
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise _lib.B200SRError("FlatAdam implements plain Adam (weight_decay=0, amsgrad=False, maximize=False)")
        return group

    @torch.no_grad()
    def host_pre_step(self):
        """Host side of a step: make sure the flat buffers are bound and advance the host mirror of the step count.
        No device work and no host->device traffic: the kernel keeps its own count."""
        engine = self._model._get_engine()
        engine.ensure_ready(next(self._model.parameters()).device)  # flat parameter storage exists from here on
        if self._bound_to != engine.flat_p.data_ptr():
            self._bind(engine)
        self._check_group()
        self._nsteps += 1

    @torch.no_grad()
    def device_step(self, grad_scale: float = 1.0):
        """Device side of a step (graph-capturable): one kernel over the flat buffers."""
        engine = self._model._get_engine()
        group = self._check_group()
        b1, b2 = group["betas"]
        call("b200sr_adam_step_auto", ptr(engine.flat_p), ptr(engine.flat_g), ptr(self._m), ptr(self._v),
             engine.p_total, float(group["lr"]), float(b1), float(b2), float(group["eps"]), ptr(self._step_dev),
             float(grad_scale), _lib.current_stream_ptr())
        engine.mark_weights_dirty()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """Consumes engine.flat_g (NOT p.grad): the engine's backward is the only producer of gradients."""
        if closure is not None:
            raise _lib.B200SRError("FlatAdam does not support closures")
        self.host_pre_step()
        self.device_step(grad_scale)

    # ---- checkpoint layout of torch.optim.Adam ---------------------------------------------------------------------
    def state_dict(self):
        for st in self.state.values():
            if "step" in st:
                st["step"] = torch.tensor(float(self._nsteps), dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """torch replaces self.state with fresh tensors: drop the binding so that the next step copies the loaded
        moments and step count into the flat buffers (also when a step has already been taken)."""
        super().load_state_dict(state_dict)
        self._bound_to = None
        self._nsteps = 0

    def __setstate__(self, state):
        super().__setstate__(state)
        self._bound_to = None
