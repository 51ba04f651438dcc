"""Fast-DDPM conditional denoiser on the b200sr kernels — SURVEY.md §8(f) row 4, BASELINE configs[4].

Mirror of the reference registry version (/root/reference/src/ModelLoader.py:471-636): `sinusoidal_timestep_embedding`,
`FastNoiseScheduler` (10 of the 1000 linear-beta steps, 40/60 split at t=699, :497-513), `DoubleConv`, `UNet2D(in_ch=3,
base_ch=64, time_dim=256)` and `FastDDPM(T=10, device)` with the same constructor signatures, module tree and
state_dict layout (26 entries under `unet.`, 2,162,177 parameters). `FastDDPM.forward(cond, target, t)` is the
noise-prediction MSE (:595-602); `FastDDPM.sample(cond, device)` the deterministic 10-step DDIM sampler (:604-636).

Engine (B200-first, not a translation):
  * the 256 time-embedding channels the reference tiles over every pixel and feeds through the 259-channel first conv
    (:565-570) are folded into a per-sample, per-border-class bias table (csrc/fastddpm.cuh): 19.3 of the 19.55 GFLOP
    per sample of that layer are never issued, the result is the same sum re-associated; its backward needs only
    border sums of dz;
  * q_sample (:597-599) is fused into the first conv's input load — x_t is never materialised;
  * every other 3x3 conv (+bias +ReLU epilogue) runs on the tcgen05 persistent implicit-GEMM kernel of the UNet path,
    weight gradients on wgrad3x3 (the 192-channel concat of up1 is split 128 + 64 over two launches);
  * nearest-2x upsampling writes straight into the channel slot of the decoder's concat buffer (no torch.cat),
    the encoder's second conv writes the other slot;
  * ReLU backward is fused with the bias-gradient reduction; one flat Adam launch; optional global-norm clipping.
Training shards the batch across ranks with one gradient all-reduce (ddp.BucketReducer); sampling shards the batch
with no collective.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr
from .engine import PACK_CONV_DGRAD, PACK_CONV_FWD, UNPACK_CONV_WGRAD, _PACK_JOB_DTYPE, _align, _jobs_to_device

_BIAS_JOB_DTYPE = np.dtype([("ps", "<u8"), ("dst", "<u8"), ("C", "<i4"), ("rows", "<i4"), ("stride", "<i4"),
                            ("pad", "<i4")])
_EPI_STATS_REPLICAS = 4  # statistic replicas of a conv epilogue (bias sums of the layers masked inside a dgrad)
# layers whose incoming gradient is produced by the dgrad of the conv above them: mask + bias sums in its epilogue
_MASKED_IN_DGRAD = {"up1.0": "up1.2", "up2.0": "up2.2", "down2.0": "down2.2", "down1.0": "down1.2"}


def sinusoidal_timestep_embedding(timesteps, dim):
    """Reference ModelLoader.py:475-487 (host-side restatement used by tests; the engine computes it in-kernel)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


class FastNoiseScheduler:
    """Reference ModelLoader.py:490-518: T of the 1000 linear-beta DDPM steps, denser late in the chain."""

    def __init__(self, T, device):
        self.T = T
        self.device = device
        beta = torch.linspace(1e-4, 0.02, 1000)
        alpha = 1.0 - beta
        alpha_bar = torch.cumprod(alpha, 0)
        boundary = 699
        late_steps = int(T * 0.6)
        early_steps = T - late_steps
        idx_early = torch.linspace(0, boundary, early_steps).long()
        idx_late = torch.linspace(boundary, 999, late_steps).long()
        idxs = torch.sort(torch.cat([idx_early, idx_late]))[0]
        self.idxs = idxs
        self.beta = beta[idxs].to(device)
        self.alpha = alpha[idxs].to(device)
        self.alpha_bar = alpha_bar[idxs].to(device)
        self._alpha_bar_host = [float(v) for v in alpha_bar[idxs]]
        # (sqrt(a_bar), sqrt(1 - a_bar)) per step: the q_sample coefficients the first-conv kernel gathers per sample
        ab = alpha_bar[idxs]
        self.coef_table = torch.stack([torch.sqrt(ab), torch.sqrt(1 - ab)], dim=1).contiguous().to(device)

    def to(self, device):
        self.device = device
        self.beta, self.alpha, self.alpha_bar = self.beta.to(device), self.alpha.to(device), self.alpha_bar.to(device)
        self.coef_table = self.coef_table.to(device)
        return self

    def q_sample(self, x0, t, noise):
        """Forward diffusion (:515-518) on the b200sr kernel."""
        if not x0.is_cuda:
            raise _lib.B200SRError("b200sr FastNoiseScheduler.q_sample runs on CUDA only; there is no CPU path")
        x0 = x0.contiguous().float()
        noise = noise.contiguous().float()
        B, C, H, W = x0.shape
        coef = self.coef_table.to(x0.device).index_select(0, t.reshape(-1).long()).contiguous()
        out = torch.empty_like(x0)
        call("b200sr_fd_q_sample", ptr(x0), ptr(noise), ptr(coef), ptr(out), B, C * H, W, _lib.current_stream_ptr())
        return out


class DoubleConv(nn.Module):
    """Parameter container (reference ModelLoader.py:521-533): Conv3x3+bias, ReLU, Conv3x3+bias, ReLU."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.ReLU(True),
                                   nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.ReLU(True))

    def forward(self, x):
        raise _lib.B200SRError("DoubleConv is a parameter container in b200sr: call the parent UNet2D")


class _UNet2DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, t, *params):
        ctx.model = model
        return model._get_engine().forward(x[:, 0:1].contiguous(), x[:, 1:3].contiguous(), t, keep=True)

    @staticmethod
    def backward(ctx, dout):
        engine = ctx.model._get_engine()
        engine.backward(dout)
        flat = engine.flat_g.clone()
        grads = [flat[off:off + p.numel()].view(p.shape) for p, off in zip(engine._params(), engine.p_off)]
        return (None, None, None, *grads)


class UNet2D(nn.Module):
    """(B,3,H,W) [x_t, pre, post] + t (B,) -> predicted noise (B,1,H,W). Reference ModelLoader.py:536-585."""

    def __init__(self, in_ch=3, base_ch=64, time_dim=256):
        super().__init__()
        self.in_ch, self.base_ch, self.time_dim = in_ch, base_ch, time_dim
        self.time_mlp = nn.Sequential(nn.Linear(time_dim, time_dim), nn.ReLU(True), nn.Linear(time_dim, time_dim))
        self.inc = DoubleConv(in_ch + time_dim, base_ch)
        self.down1 = DoubleConv(base_ch, base_ch * 2)
        self.down2 = DoubleConv(base_ch * 2, base_ch * 4)
        self.up2 = DoubleConv(base_ch * 4 + base_ch * 2, base_ch * 2)
        self.up1 = DoubleConv(base_ch * 2 + base_ch, base_ch)
        self.outc = nn.Conv2d(base_ch, 1, 1)

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = FastDDPMEngine(self)
            self.__dict__["_engine"] = eng
        return eng

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_engine", None)
        return state

    def forward(self, x, t):
        if not x.is_cuda:
            raise _lib.B200SRError("b200sr.UNet2D runs on CUDA sm_100a only; there is no CPU/torch fallback")
        if x.dim() != 4 or x.shape[1] != 3:
            raise _lib.B200SRError(f"expected input (B,3,H,W) = [x_t, pre, post], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _UNet2DFunction.apply(self, x, t, *self.parameters())
        x = x.float()
        return self._get_engine().forward(x[:, 0:1].contiguous(), x[:, 1:3].contiguous(), t)


class _Conv:
    def __init__(self, name, mod, cin, cout, level):
        self.name, self.mod, self.cin, self.cout, self.level = name, mod, cin, cout, level


class FastDDPMEngine:
    """Issues the C-ABI ops of one UNet2D forward / backward. Contract shared with UNetEngine (FlatAdam, BucketReducer):
    flat_p / flat_g / p_off / p_total / _params() / ensure_ready() / mark_weights_dirty()."""

    def __init__(self, model):
        if model.in_ch != 3 or model.base_ch != 64 or model.time_dim != 256:
            raise NotImplementedError("b200sr Fast-DDPM engine implements the reference configuration "
                                      "UNet2D(in_ch=3, base_ch=64, time_dim=256)")
        self.model = model
        m = model
        self.convs = [
            _Conv("inc.0", m.inc.block[0], 259, 64, 0), _Conv("inc.2", m.inc.block[2], 64, 64, 0),
            _Conv("down1.0", m.down1.block[0], 64, 128, 1), _Conv("down1.2", m.down1.block[2], 128, 128, 1),
            _Conv("down2.0", m.down2.block[0], 128, 256, 2), _Conv("down2.2", m.down2.block[2], 256, 256, 2),
            _Conv("up2.0", m.up2.block[0], 384, 128, 1), _Conv("up2.2", m.up2.block[2], 128, 128, 1),
            _Conv("up1.0", m.up1.block[0], 192, 64, 0), _Conv("up1.2", m.up1.block[2], 64, 64, 0),
        ]
        self.by_name = {c.name: c for c in self.convs}
        self.device = None
        self.flat_p = None
        self._plans = {}
        self._packed_version = None
        self._saved = None
        self.fuse_relu_bwd = os.environ.get("B200SR_FD_UNFUSED") is None

    # ---- flat parameter storage ---------------------------------------------------------------------------------
    def _params(self):
        return list(self.model.parameters())

    def mark_weights_dirty(self):
        self._packed_version = None

    def _is_flat(self):
        if self.flat_p is None:
            return False
        base = self.flat_p.data_ptr()
        return all(p.data.data_ptr() == base + 4 * off for p, off in zip(self._params(), self.p_off))

    def ensure_ready(self, device):
        if self.device == device and self._is_flat():
            return
        _lib.require_device()
        self.device = device
        params = self._params()
        for p in params:
            if p.device != device or p.dtype != torch.float32:
                raise _lib.B200SRError(f"all UNet2D parameters must be fp32 on {device} (got {p.dtype} on {p.device})")
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        self.p_off, self.p_total = offs, total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_G = torch.zeros(total, dtype=torch.float32, device=device)  # wgrad workspace, kernel layouts
        self.grad_views = []
        for p, off in zip(params, offs):
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grad_views.append(self.flat_g[off:off + p.numel()].view(p.shape))
        self.off_of = {id(p): off for p, off in zip(params, offs)}

        tc = self.convs[1:]  # tensor-core convs (everything but the 259-channel first conv)
        wp_total = 0
        self.wp_fwd, self.wp_dgrad = {}, {}
        for c in tc:
            n = c.cout * c.cin * 9
            self.wp_fwd[c.name] = wp_total
            wp_total += _align(n)
            self.wp_dgrad[c.name] = wp_total
            wp_total += _align(n)
        self.flat_wp = torch.zeros(wp_total, dtype=torch.bfloat16, device=device)
        pack = np.zeros(2 * len(tc), dtype=_PACK_JOB_DTYPE)
        unpack = []
        wpb, gb, Gb = self.flat_wp.data_ptr(), self.flat_g.data_ptr(), self.flat_G.data_ptr()
        for i, c in enumerate(tc):
            w = c.mod.weight
            n = w.numel()
            pack[2 * i] = (w.data_ptr(), wpb + 2 * self.wp_fwd[c.name], PACK_CONV_FWD, c.cout, c.cin, 0, n)
            pack[2 * i + 1] = (w.data_ptr(), wpb + 2 * self.wp_dgrad[c.name], PACK_CONV_DGRAD, c.cout, c.cin, 0, n)
            off = self.off_of[id(w)]
            if c.cin == 192:
                # wgrad3x3 tiles Cin as 64 or multiples of 128: two launches over channel ranges [0,128) and [128,192),
                # each with its own G[9][cin_part][cout] workspace, unpacked into the (cout,192,3,3) gradient
                n0 = 9 * 128 * c.cout
                unpack.append((Gb + 4 * off, gb + 4 * off, UNPACK_CONV_WGRAD, c.cout, 128, 192, n0))
                unpack.append((Gb + 4 * (off + n0), gb + 4 * (off + 128 * 9), UNPACK_CONV_WGRAD, c.cout, 64, 192, n - n0))
            else:
                unpack.append((Gb + 4 * off, gb + 4 * off, UNPACK_CONV_WGRAD, c.cout, c.cin, 0, n))
        unpack_np = np.zeros(len(unpack), dtype=_PACK_JOB_DTYPE)
        for i, u in enumerate(unpack):
            unpack_np[i] = u
        self.pack_jobs, self.n_pack = _jobs_to_device(pack, device), len(pack)
        self.unpack_jobs, self.n_unpack = _jobs_to_device(unpack_np, device), len(unpack_np)
        self._plans = {}
        self._packed_version = None

    def _wp(self, table, name):
        return self.flat_wp.data_ptr() + 2 * table[name]

    def _G(self, p, elem_off=0):
        return self.flat_G.data_ptr() + 4 * (self.off_of[id(p)] + elem_off)

    def _g(self, p):
        return self.flat_g.data_ptr() + 4 * self.off_of[id(p)]

    def _state_version(self):
        return sum(p._version for p in self.model.parameters())

    def _repack_if_needed(self, st):
        ver = self._state_version()
        if ver != self._packed_version:
            call("b200sr_pack_jobs", self.pack_jobs.data_ptr(), self.n_pack, st)
            self._packed_version = ver

    # ---- activation plan ----------------------------------------------------------------------------------------
    def _plan(self, B, H, W):
        key = (B, H, W)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        if H % 16 != 0 or W % 16 != 0:
            # the first layer works on 16x16-pixel tiles; the deeper levels may be ragged (csrc/conv3x3.cuh)
            raise _lib.B200SRError(f"b200sr UNet2D needs H % 16 == 0 and W % 16 == 0 (got {H}x{W})")
        dev, bf, f32 = self.device, torch.bfloat16, torch.float32
        H1, W1, H2, W2 = H // 2, W // 2, H // 4, W // 4

        def buf(h, w, c):
            return torch.empty((B, h, w, c), dtype=bf, device=dev)

        plan = {"B": B, "H": H, "W": W,
                "a_inc": buf(H, W, 64), "cat1": buf(H, W, 192), "p1": buf(H1, W1, 64),
                "a_d1": buf(H1, W1, 128), "cat2": buf(H1, W1, 384), "p2": buf(H2, W2, 128),
                "a_d2": buf(H2, W2, 256), "c3": buf(H2, W2, 256),
                "a_u2": buf(H1, W1, 128), "u2": buf(H1, W1, 128),
                "a_u1": buf(H, W, 64), "u1": buf(H, W, 64),
                "emb": torch.empty(B, 256, dtype=f32, device=dev), "hid": torch.empty(B, 256, dtype=f32, device=dev),
                "e": torch.empty(B, 256, dtype=f32, device=dev), "tb": torch.empty(B, 9, 64, dtype=f32, device=dev)}
        self._plans[key] = plan
        return plan

    def _bwd_plan(self, plan):
        if "g0" in plan:
            return
        B, H, W = plan["B"], plan["H"], plan["W"]
        dev, bf, f32 = self.device, torch.bfloat16, torch.float32
        H1, W1, H2, W2 = H // 2, W // 2, H // 4, W // 4
        # two ping-pong gradient buffers sized for the largest dense tensor (64 ch at full resolution)
        n = B * H * W * 64
        plan["g0"] = torch.empty(n, dtype=bf, device=dev)
        plan["g1"] = torch.empty(n, dtype=bf, device=dev)
        plan["d_cat1"] = torch.empty((B, H, W, 192), dtype=bf, device=dev)
        plan["d_cat2"] = torch.empty((B, H1, W1, 384), dtype=bf, device=dev)
        ps_total, plan["ps_off"] = 0, {}
        for c in self.convs:
            plan["ps_off"][c.name] = ps_total
            ps_total += B * c.cout
        plan["ps"] = torch.zeros(ps_total, dtype=f32, device=dev)
        # bias sums of the layers masked inside a dgrad epilogue arrive as [replicas][2][C] statistics
        st_total, plan["st_off"] = 0, {}
        for name in _MASKED_IN_DGRAD:
            plan["st_off"][name] = st_total
            st_total += _EPI_STATS_REPLICAS * 2 * self.by_name[name].cout
        plan["epi_stats"] = torch.zeros(st_total, dtype=f32, device=dev)
        jobs = np.zeros(len(self.convs), dtype=_BIAS_JOB_DTYPE)
        for i, c in enumerate(self.convs):
            if self.fuse_relu_bwd and c.name in _MASKED_IN_DGRAD:
                jobs[i] = (plan["epi_stats"].data_ptr() + 4 * plan["st_off"][c.name], self._g(c.mod.bias), c.cout,
                           _EPI_STATS_REPLICAS, 2 * c.cout, 0)
            else:
                jobs[i] = (plan["ps"].data_ptr() + 4 * plan["ps_off"][c.name], self._g(c.mod.bias), c.cout, 0, 0, 0)
        plan["bias_jobs"] = _jobs_to_device(jobs, dev)
        plan["S"] = torch.empty(B, 9, 64, dtype=f32, device=dev)
        plan["de"] = torch.empty(B, 256, dtype=f32, device=dev)
        plan["dh"] = torch.empty(B, 256, dtype=f32, device=dev)

    # ---- forward ------------------------------------------------------------------------------------------------
    def forward(self, x0, cond, t, noise=None, coef=None, keep=False):
        """x0: (B,1,H,W) f32 — x_t itself, or the clean target when `noise`/`coef` are given (q_sample fused into the first
        conv: x_t = coef[b,0]*x0 + coef[b,1]*noise); cond: (B,2,H,W) f32; t: (B,) int64. Returns eps (B,1,H,W) f32."""
        if x0.dim() != 4 or x0.shape[1] != 1 or cond.dim() != 4 or cond.shape[1] != 2 or x0.shape[0] != cond.shape[0] \
                or x0.shape[2:] != cond.shape[2:]:
            raise _lib.B200SRError(f"expected x (B,1,H,W) and cond (B,2,H,W), got {tuple(x0.shape)} / {tuple(cond.shape)}")
        if (noise is None) != (coef is None):
            raise _lib.B200SRError("noise and coef go together")
        dev = x0.device
        self.ensure_ready(dev)
        x0 = x0.contiguous().float()
        cond = cond.contiguous().float()
        t = t.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        B, _, H, W = x0.shape
        if t.numel() != B:
            raise _lib.B200SRError(f"t must have one entry per sample ({B}), got {t.numel()}")
        if noise is not None:
            noise = noise.contiguous().float()
            coef = coef.contiguous().float()
        plan = self._plan(B, H, W)
        st = _lib.current_stream_ptr()
        m = self.model
        self._repack_if_needed(st)
        H1, W1, H2, W2 = H // 2, W // 2, H // 4, W // 4
        c = self.by_name

        def conv(name, src, s_stride, s_off, dst, d_stride, d_off, h, w):
            cv = c[name]
            call("b200sr_conv3x3_fwd", ptr(src), s_stride, s_off, cv.cin, self._wp(self.wp_fwd, name), cv.cout, B, h, w,
                 ptr(dst), d_stride, d_off, None, ptr(cv.mod.bias), 1, None, 0, st)

        l1, l2 = m.time_mlp[0], m.time_mlp[2]
        call("b200sr_fd_time_mlp_fwd", ptr(t), ptr(l1.weight), ptr(l1.bias), ptr(l2.weight), ptr(l2.bias),
             ptr(plan["emb"]), ptr(plan["hid"]), ptr(plan["e"]), B, st)
        w_in = c["inc.0"].mod
        call("b200sr_fd_time_bias", ptr(plan["e"]), ptr(w_in.weight), ptr(w_in.bias), ptr(plan["tb"]), B, st)
        call("b200sr_fd_convin_fwd", ptr(x0), ptr(noise), ptr(coef), ptr(cond), ptr(w_in.weight), ptr(plan["tb"]),
             ptr(plan["a_inc"]), B, H, W, st)
        conv("inc.2", plan["a_inc"], 64, 0, plan["cat1"], 192, 128, H, W)                 # c1 -> cat1[128:192]
        call("b200sr_maxpool2x2_fwd", ptr(plan["cat1"]), 192, 128, 64, ptr(plan["p1"]), B, H, W, st)
        conv("down1.0", plan["p1"], 64, 0, plan["a_d1"], 128, 0, H1, W1)
        conv("down1.2", plan["a_d1"], 128, 0, plan["cat2"], 384, 256, H1, W1)              # c2 -> cat2[256:384]
        call("b200sr_maxpool2x2_fwd", ptr(plan["cat2"]), 384, 256, 128, ptr(plan["p2"]), B, H1, W1, st)
        conv("down2.0", plan["p2"], 128, 0, plan["a_d2"], 256, 0, H2, W2)
        conv("down2.2", plan["a_d2"], 256, 0, plan["c3"], 256, 0, H2, W2)
        call("b200sr_fd_upsample2x_fwd", ptr(plan["c3"]), 256, ptr(plan["cat2"]), 384, 0, B, H2, W2, st)
        conv("up2.0", plan["cat2"], 384, 0, plan["a_u2"], 128, 0, H1, W1)
        conv("up2.2", plan["a_u2"], 128, 0, plan["u2"], 128, 0, H1, W1)
        call("b200sr_fd_upsample2x_fwd", ptr(plan["u2"]), 128, ptr(plan["cat1"]), 192, 0, B, H1, W1, st)
        conv("up1.0", plan["cat1"], 192, 0, plan["a_u1"], 64, 0, H, W)
        conv("up1.2", plan["a_u1"], 64, 0, plan["u1"], 64, 0, H, W)
        y = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        call("b200sr_head_fwd", ptr(plan["u1"]), ptr(m.outc.weight), ptr(m.outc.bias), ptr(y), B * H * W, st)
        if keep:
            self._saved = (plan, x0, cond, noise, coef)
        return y

    # ---- backward -----------------------------------------------------------------------------------------------
    def backward(self, dout, bucket_hook=None):
        """Gradients of every parameter into flat_g (zeroed here). dout: (B,1,H,W) f32."""
        if self._saved is None:
            raise _lib.B200SRError("backward() without a preceding forward(keep=True)")
        plan, x0, cond, noise, coef = self._saved
        self._bwd_plan(plan)
        B, H, W = plan["B"], plan["H"], plan["W"]
        H1, W1, H2, W2 = H // 2, W // 2, H // 4, W // 4
        st = _lib.current_stream_ptr()
        m = self.model
        c = self.by_name
        dout = dout.contiguous().float()
        self.flat_g.zero_()
        self.flat_G.zero_()
        plan["ps"].zero_()
        plan["epi_stats"].zero_()
        g0, g1 = plan["g0"].data_ptr(), plan["g1"].data_ptr()
        ps = plan["ps"].data_ptr()

        def relu_bwd(name, dy, dy_stride, dy_off, act, a_stride, a_off, dz, h, w):
            cv = c[name]
            call("b200sr_fd_relu_bwd_bias", dy, dy_stride, dy_off, ptr(act), a_stride, a_off, dz,
                 ps + 4 * plan["ps_off"][name], cv.cout, B, h, w, st)

        def wgrad(name, x, x_stride, x_off, dz, h, w):
            cv = c[name]
            wt = cv.mod.weight
            if cv.cin == 192:
                call("b200sr_conv3x3_wgrad", ptr(x), x_stride, x_off, 128, dz, cv.cout, 0, cv.cout, B, h, w,
                     self._G(wt), st)
                call("b200sr_conv3x3_wgrad", ptr(x), x_stride, x_off + 128, 64, dz, cv.cout, 0, cv.cout, B, h, w,
                     self._G(wt, 9 * 128 * cv.cout), st)
            else:
                call("b200sr_conv3x3_wgrad", ptr(x), x_stride, x_off, cv.cin, dz, cv.cout, 0, cv.cout, B, h, w,
                     self._G(wt), st)

        def dgrad(name, dz, dx, dx_stride, h, w):
            cv = c[name]
            call("b200sr_conv3x3_dgrad", dz, cv.cout, 0, cv.cout, self._wp(self.wp_dgrad, name), cv.cin, B, h, w, dx,
                 dx_stride, 0, None, 0, st)

        def dgrad_into(below, act, dz, dx, h, w):
            """dgrad of the conv above `below`, producing dz of `below` directly: with the fusion on, the ReLU mask of
            `below` (its stored activation) and its bias sums are applied in the dgrad epilogue; otherwise dgrad followed
            by the separate ReLU-backward + bias pass (in place)."""
            above = _MASKED_IN_DGRAD[below]
            cv = c[above]
            if self.fuse_relu_bwd:
                call("b200sr_conv3x3_dgrad_relu", dz, cv.cout, 0, cv.cout, self._wp(self.wp_dgrad, above), cv.cin, B, h, w,
                     dx, cv.cin, 0, ptr(act), cv.cin, 0, plan["epi_stats"].data_ptr() + 4 * plan["st_off"][below],
                     _EPI_STATS_REPLICAS, st)
            else:
                dgrad(above, dz, dx, cv.cin, h, w)
                relu_bwd(below, dx, cv.cin, 0, act, cv.cin, 0, dx, h, w)

        fused = self.fuse_relu_bwd  # ReLU mask + bias sums applied where a bandwidth-bound kernel forms the gradient

        def ps_of(name):
            return ps + 4 * plan["ps_off"][name]

        if fused:
            call("b200sr_fd_head_bwd_relu", ptr(dout), ptr(plan["u1"]), ptr(m.outc.weight), g0, self._g(m.outc.weight),
                 self._g(m.outc.bias), ps_of("up1.2"), B, H, W, st)
        else:
            call("b200sr_head_bwd", ptr(dout), ptr(plan["u1"]), ptr(m.outc.weight), g0, self._g(m.outc.weight),
                 self._g(m.outc.bias), B * H * W, st)
            relu_bwd("up1.2", g0, 64, 0, plan["u1"], 64, 0, g0, H, W)
        # up1
        wgrad("up1.2", plan["a_u1"], 64, 0, g0, H, W)
        dgrad_into("up1.0", plan["a_u1"], g0, g1, H, W)
        wgrad("up1.0", plan["cat1"], 192, 0, g1, H, W)
        dgrad("up1.0", g1, ptr(plan["d_cat1"]), 192, H, W)
        if fused:
            call("b200sr_fd_upsample2x_bwd_relu", ptr(plan["d_cat1"]), 192, 0, ptr(plan["u2"]), g0, ps_of("up2.2"), 128,
                 B, H1, W1, st)
        else:
            call("b200sr_fd_upsample2x_bwd", ptr(plan["d_cat1"]), 192, 0, 128, g0, B, H1, W1, st)  # -> d u2
            relu_bwd("up2.2", g0, 128, 0, plan["u2"], 128, 0, g0, H1, W1)
        # up2
        wgrad("up2.2", plan["a_u2"], 128, 0, g0, H1, W1)
        dgrad_into("up2.0", plan["a_u2"], g0, g1, H1, W1)
        wgrad("up2.0", plan["cat2"], 384, 0, g1, H1, W1)
        dgrad("up2.0", g1, ptr(plan["d_cat2"]), 384, H1, W1)
        if fused:
            call("b200sr_fd_upsample2x_bwd_relu", ptr(plan["d_cat2"]), 384, 0, ptr(plan["c3"]), g0, ps_of("down2.2"), 256,
                 B, H2, W2, st)
        else:
            call("b200sr_fd_upsample2x_bwd", ptr(plan["d_cat2"]), 384, 0, 256, g0, B, H2, W2, st)  # -> d c3
            relu_bwd("down2.2", g0, 256, 0, plan["c3"], 256, 0, g0, H2, W2)
        # down2
        wgrad("down2.2", plan["a_d2"], 256, 0, g0, H2, W2)
        dgrad_into("down2.0", plan["a_d2"], g0, g1, H2, W2)
        wgrad("down2.0", plan["p2"], 128, 0, g1, H2, W2)
        dgrad("down2.0", g1, g0, 128, H2, W2)                                                      # -> d p2
        if fused:
            call("b200sr_fd_maxpool2x2_bwd_relu", ptr(plan["cat2"]), 384, 256, g0, ptr(plan["d_cat2"]), 384, 256, 128, g1,
                 ps_of("down1.2"), B, H1, W1, st)
        else:
            call("b200sr_maxpool2x2_bwd", ptr(plan["cat2"]), 384, 256, g0, ptr(plan["d_cat2"]), 384, 256, 128, g1, B, H1,
                 W1, st)                                                                           # -> d c2 (+ skip)
            relu_bwd("down1.2", g1, 128, 0, plan["cat2"], 384, 256, g1, H1, W1)
        # down1
        wgrad("down1.2", plan["a_d1"], 128, 0, g1, H1, W1)
        dgrad_into("down1.0", plan["a_d1"], g1, g0, H1, W1)
        wgrad("down1.0", plan["p1"], 64, 0, g0, H1, W1)
        dgrad("down1.0", g0, g1, 64, H1, W1)                                                       # -> d p1
        if fused:
            call("b200sr_fd_maxpool2x2_bwd_relu", ptr(plan["cat1"]), 192, 128, g1, ptr(plan["d_cat1"]), 192, 128, 64, g0,
                 ps_of("inc.2"), B, H, W, st)
        else:
            call("b200sr_maxpool2x2_bwd", ptr(plan["cat1"]), 192, 128, g1, ptr(plan["d_cat1"]), 192, 128, 64, g0, B, H, W,
                 st)                                                                               # -> d c1 (+ skip)
            relu_bwd("inc.2", g0, 64, 0, plan["cat1"], 192, 128, g0, H, W)
        # inc
        wgrad("inc.2", plan["a_inc"], 64, 0, g0, H, W)
        dgrad("inc.2", g0, g1, 64, H, W)
        relu_bwd("inc.0", g1, 64, 0, plan["a_inc"], 64, 0, g1, H, W)
        w_in = c["inc.0"].mod
        call("b200sr_fd_convin_wgrad", ptr(x0), ptr(noise), ptr(coef), ptr(cond), g1, self._g(w_in.weight), B, H, W, st)
        l1, l2 = m.time_mlp[0], m.time_mlp[2]
        call("b200sr_fd_time_bwd", g1, ps + 4 * plan["ps_off"]["inc.0"], ptr(plan["e"]), ptr(plan["emb"]),
             ptr(plan["hid"]), ptr(w_in.weight), ptr(l2.weight), ptr(plan["S"]), ptr(plan["de"]), ptr(plan["dh"]),
             self._g(w_in.weight), self._g(l1.weight), self._g(l1.bias), self._g(l2.weight), self._g(l2.bias), B, H, W, st)
        call("b200sr_pack_jobs", self.unpack_jobs.data_ptr(), self.n_unpack, st)
        call("b200sr_fd_bias_finish", plan["bias_jobs"].data_ptr(), len(self.convs), B, st)
        if bucket_hook is not None:
            bucket_hook(0, self.p_total)
        self._saved = None
        return self.grad_views


class FastDDPM(nn.Module):
    """Reference ModelLoader.py:588-636. `forward(cond, target, t)` = noise-prediction MSE; `sample(cond, device)`."""

    def __init__(self, T=10, device='cuda'):
        super().__init__()
        self.device = device
        self.scheduler = FastNoiseScheduler(T, device)
        self.unet = UNet2D(in_ch=3, base_ch=64, time_dim=256).to(device)
        from .losses import CombinedLoss
        self.__dict__["_mse"] = CombinedLoss(mse_weight=1.0, ssim_weight=0.0)
        self.__dict__["sample_graphs"] = os.environ.get("B200SR_NO_SAMPLE_GRAPH") is None
        self.__dict__["_sample_graph_cache"], self.__dict__["_sample_graph_calls"] = {}, {}

    def _coef(self, t, dev):
        sch = self.scheduler
        if sch.coef_table.device != dev:
            sch.to(dev)
        return sch.coef_table.index_select(0, t.reshape(-1).to(dev).long()).contiguous()

    def loss_and_grad(self, cond, target, t, noise=None, need_grad=True):
        """One denoiser evaluation on a noised target: returns (loss 0-d f32 device tensor, d loss / d eps_pred or None).
        `noise` (B,1,H,W) may be supplied (parity tests); default torch.randn_like(target) as the reference (:597)."""
        if not target.is_cuda:
            raise _lib.B200SRError("b200sr.FastDDPM runs on CUDA sm_100a only; there is no CPU/torch fallback")
        target = target.contiguous().float()
        if noise is None:
            noise = torch.randn_like(target)
        coef = self._coef(t, target.device)
        eps = self.unet._get_engine().forward(target, cond, t, noise=noise, coef=coef, keep=need_grad)
        return self._mse.value_and_grad(eps, noise, need_grad=need_grad)

    def forward(self, cond, target, t, noise=None):
        """Noise-prediction MSE (reference ModelLoader.py:595-602). Like the reference's, the returned loss is
        differentiable: `loss.backward()` runs the engine's hand-written backward and fills `p.grad` of the denoiser's
        parameters (an autograd node around forward + backward). Under no_grad / frozen parameters it is a plain value."""
        params = list(self.unet.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _FastDDPMLossFn.apply(self, cond, target, t, noise, *params)
        return self.loss_and_grad(cond, target, t, noise=noise, need_grad=False)[0]

    @torch.no_grad()
    def sample(self, cond, device=None, noise=None):
        """Deterministic DDIM sampling (:604-636): T denoiser evaluations, x updated in place, clamp(-1,1).
        `noise` (B,1,H,W): the initial x_T (default torch.randn, as the reference draws it)."""
        if not cond.is_cuda:
            raise _lib.B200SRError("b200sr.FastDDPM.sample runs on CUDA sm_100a only; there is no CPU/torch fallback")
        dev = cond.device
        cond = cond.contiguous().float()
        B, _, H, W = cond.shape
        x = (torch.randn(B, 1, H, W, device=dev) if noise is None else noise.to(dev).float().clone()).contiguous()
        engine = self.unet._get_engine()
        if not self.sample_graphs or torch.cuda.is_current_stream_capturing():
            return self._sample_loop(engine, x, cond)
        # the whole T-step chain (T x ~22 launches + T DDIM updates) is replayed from one CUDA graph per input shape
        key = (B, H, W, dev)
        g = self._sample_graph_cache.get(key)
        if g is not None and g[3] != engine.flat_p.data_ptr():
            g = None  # parameters were re-flattened (model.to(), new tensors): the graph points into freed buffers
        if g is None:
            n = self._sample_graph_calls.get(key, 0)
            self._sample_graph_calls[key] = n + 1
            if n < 2:  # lazy set-up (plan buffers, TMA descriptors, packed weights) happens eagerly first
                return self._sample_loop(engine, x, cond)
            static_x, static_cond = x.clone(), cond.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._sample_loop(engine, static_x, static_cond)
            g = (graph, static_x, static_cond, engine.flat_p.data_ptr())
            self._sample_graph_cache[key] = g
        graph, static_x, static_cond, _ = g
        engine.ensure_ready(dev)
        engine._repack_if_needed(_lib.current_stream_ptr())  # packed bf16 weights follow the parameters
        static_x.copy_(x)
        static_cond.copy_(cond)
        graph.replay()
        return static_x.clone()

    def _sample_loop(self, engine, x, cond):
        """T denoiser evaluations + DDIM updates, in place on x (which is also returned)."""
        dev = x.device
        B, _, H, W = x.shape
        T = self.scheduler.T
        ab = self.scheduler._alpha_bar_host
        n = B * H * W
        for i in reversed(range(T)):
            t = torch.full((B,), i, device=dev, dtype=torch.long)
            eps = engine.forward(x, cond, t)
            a_prev = ab[i - 1] if i > 0 else 1.0
            call("b200sr_fd_ddim_update", ptr(x), ptr(eps), ab[i], a_prev, 1 if i == 0 else 0, n,
                 _lib.current_stream_ptr())
        return x


class _FastDDPMLossFn(torch.autograd.Function):
    """Autograd node of FastDDPM.forward: engine forward + fused MSE in forward(), engine backward in backward()."""

    @staticmethod
    def forward(ctx, model, cond, target, t, noise, *params):
        loss, dout = model.loss_and_grad(cond, target, t, noise=noise, need_grad=True)
        ctx.model = model
        ctx.save_for_backward(dout)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (dout,) = ctx.saved_tensors
        engine = ctx.model.unet._get_engine()
        engine.backward(dout * gout)
        flat = engine.flat_g.clone()  # independent copy: the flat gradient buffer is reused by the next step
        grads = [flat[off:off + p.numel()].view(p.shape) for p, off in zip(engine._params(), engine.p_off)]
        return (None, None, None, None, None, *grads)


class FastDDPMTrainer:
    """Denoiser train step: t ~ U{0..T-1}, noise ~ N(0,1), loss = MSE(eps_pred, noise), global-norm clip, Adam.
    Hyper-parameters: FastDDPM_Training_Fixed.ipynb cells 9-11 (lr 2e-4, grad-clip 1.0); the notebook of the registry
    model (FastDDPM_Simple.ipynb) is missing from the snapshot, so the step is frozen here."""

    def __init__(self, model, device="cuda", learning_rate=2e-4, grad_clip=1.0, model_save_dir="models", verbose=True):
        from .optim import FlatAdam
        self.model = model.to(device)
        self.model.scheduler.to(device)
        self.device = device
        self.grad_clip = grad_clip
        self.optimizer = FlatAdam(self.model.unet, lr=learning_rate)
        self.model_save_dir = Path(model_save_dir)
        self.model_save_dir.mkdir(parents=True, exist_ok=True)
        self._reducer = None
        self._sumsq = None
        from .ddp import broadcast_module_state, is_distributed
        if is_distributed():
            broadcast_module_state(self.model)  # every rank starts from rank 0's weights
        if verbose:
            print(f"Total parameters: {sum(p.numel() for p in self.model.parameters()):,}")

    def train_step(self, cond, target, t=None, noise=None):
        self.model.train()
        engine = self.model.unet._get_engine()
        self.optimizer.host_pre_step()
        B = target.shape[0]
        if t is None:
            t = torch.randint(0, self.model.scheduler.T, (B,), device=target.device)
        loss, dout = self.model.loss_and_grad(cond, target, t, noise=noise, need_grad=True)
        hook, scale = None, 1.0
        from .ddp import is_distributed
        if is_distributed():
            if self._reducer is None or self._reducer.flat is not engine.flat_g:
                from .ddp import BucketReducer
                self._reducer = BucketReducer(engine.flat_g)
            hook, scale = self._reducer.reduce_range, 1.0 / self._reducer.world_size
        engine.backward(dout, bucket_hook=hook)
        if hook is not None:
            self._reducer.wait()
        if self.grad_clip:
            if self._sumsq is None:
                self._sumsq = torch.zeros(1, dtype=torch.float64, device=engine.flat_g.device)
            call("b200sr_grad_clip", ptr(engine.flat_g), engine.p_total, ptr(self._sumsq), float(self.grad_clip),
                 float(scale), _lib.current_stream_ptr())
        self.optimizer.device_step(grad_scale=scale)
        return loss

    def save_checkpoint(self, epoch, val_loss, is_best=False):
        ck = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "val_loss": val_loss}
        if is_best:
            torch.save(ck, self.model_save_dir / "fastddpm_advanced_best.pth")
        torch.save(ck, self.model_save_dir / "fastddpm_latest.pth")
