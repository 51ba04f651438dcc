"""Adam over the engine's flat parameter / gradient buffers: one kernel launch for all 31 M parameters.

Reference: `optim.Adam(self.model.parameters(), lr=learning_rate)` in UNetTrainer.__init__ (unet_model.py:155) and
`self.optimizer.step()` (:185). Subclasses torch.optim.Adam so `state_dict()` / `load_state_dict()` keep the
reference checkpoint layout ('optimizer_state_dict', unet_model.py:252): the per-parameter `exp_avg` /
`exp_avg_sq` state tensors are views into two flat fp32 buffers that the kernel updates.
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
        self._bc_host = None
        self._bc_dev = None

    def _bind(self, engine):
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
        self._m, self._v = new_m, new_v
        self._nsteps = steps
        self._bound_to = engine.flat_p.data_ptr()

    def _check_group(self):
        group = self.param_groups[0]
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise _lib.B200SRError("FlatAdam implements plain Adam (weight_decay=0, amsgrad=False, maximize=False)")
        return group

    @torch.no_grad()
    def host_pre_step(self):
        """Host side of a step: advance the step count and publish the bias corrections to device memory (a 8-byte
        async copy from pinned memory). Kept apart from the kernel launch so the launch can be replayed by a graph."""
        engine = self._model._get_engine()
        engine.ensure_ready(next(self._model.parameters()).device)  # flat parameter storage exists from here on
        if self._bound_to != engine.flat_p.data_ptr():
            self._bind(engine)
        group = self._check_group()
        self._nsteps += 1
        b1, b2 = group["betas"]
        if self._bc_host is None:
            self._bc_host = torch.zeros(2, dtype=torch.float32).pin_memory()
            self._bc_dev = torch.zeros(2, dtype=torch.float32, device=engine.flat_p.device)
        self._bc_host[0] = 1.0 - b1 ** self._nsteps
        self._bc_host[1] = (1.0 - b2 ** self._nsteps) ** 0.5
        self._bc_dev.copy_(self._bc_host, non_blocking=True)
        torch._foreach_add_([self.state[p]["step"] for p in engine._params()], 1.0)

    @torch.no_grad()
    def device_step(self, grad_scale: float = 1.0):
        """Device side of a step (graph-capturable): one kernel over the flat buffers."""
        engine = self._model._get_engine()
        group = self._check_group()
        b1, b2 = group["betas"]
        call("b200sr_adam_step_dev", ptr(engine.flat_p), ptr(engine.flat_g), ptr(self._m), ptr(self._v),
             engine.p_total, float(group["lr"]), float(b1), float(b2), float(group["eps"]), ptr(self._bc_dev),
             float(grad_scale), _lib.current_stream_ptr())
        engine.mark_weights_dirty()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """Consumes engine.flat_g (NOT p.grad): the engine's backward is the only producer of gradients."""
        if closure is not None:
            raise _lib.B200SRError("FlatAdam does not support closures")
        self.host_pre_step()
        self.device_step(grad_scale)
