"""DeepCNN residual baseline on the b200sr kernels — SURVEY.md §8(f) row 3, BASELINE configs[1].

Mirror of the reference `ResidualBlock` / `DeepCNN` (/root/reference/src/ModelLoader.py:276-377): same constructor
signatures (`DeepCNN(in_channels=2, out_channels=1, num_blocks=[2,2,2,2], base_features=64)`), module tree, init
(kaiming-normal fan_out on every conv, BN weight 1 / bias 0) and state_dict layout (122 entries, 11,173,889
parameters). Everything runs at full 256x256 resolution (all strides are 1, :329-332); `avgpool` is declared but unused
by the reference forward (:361-377) and is kept only for the module tree.

Engine: 7x7 stem (direct kernel) -> BN+ReLU -> MaxPool3x3/s1 -> 8 residual blocks (tcgen05 conv3x3 forward/dgrad/
wgrad, 1x1 downsample conv on the same persistent kernel, fused `relu(bn2(z2) + identity)` tail, BatchNorm backward
whose ReLU mask comes from the stored block output) -> wide 1x1 head. Eval mode uses the same kernels with
scale/shift folded from the running statistics.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr
from .engine import (BN_EPS, BN_MOMENTUM, PACK_CONV_DGRAD, PACK_CONV_FWD, STATS_REPLICAS, UNPACK_CONV_WGRAD,
                     _FOLD_JOB_DTYPE, _PACK_JOB_DTYPE, _align, _jobs_to_device)

PACK_1X1_FWD, PACK_1X1_DGRAD, UNPACK_1X1_WGRAD = 6, 7, 8


class ResidualBlock(nn.Module):
    """Parameter container (reference ModelLoader.py:276-307)."""

    def __init__(self, in_channels, out_channels, stride=1, downsample=None):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("b200sr implements the reference DeepCNN, whose blocks all use stride 1")
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.downsample = downsample

    def forward(self, x):
        raise _lib.B200SRError("ResidualBlock is a parameter container in b200sr: call the parent DeepCNN")


class _DeepCNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        ctx.model = model
        return model._get_engine().forward(x, training=True)

    @staticmethod
    def backward(ctx, dout):
        engine = ctx.model._get_engine()
        engine.backward(dout)
        flat = engine.flat_g.clone()
        grads = [flat[off:off + p.numel()].view(p.shape) for p, off in zip(engine._params(), engine.p_off)]
        return (None, None, *grads)


class DeepCNN(nn.Module):
    """(B,2,H,W) -> (B,1,H,W). Reference ModelLoader.py:310-377."""

    def __init__(self, in_channels=2, out_channels=1, num_blocks=[2, 2, 2, 2], base_features=64):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_blocks = list(num_blocks)
        self.base_features = base_features
        f = base_features
        self.conv1 = nn.Conv2d(in_channels, f, kernel_size=7, stride=1, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(f)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=1, padding=1)
        self.layer1 = self._make_layer(f, f, num_blocks[0])
        self.layer2 = self._make_layer(f, f * 2, num_blocks[1])
        self.layer3 = self._make_layer(f * 2, f * 4, num_blocks[2])
        self.layer4 = self._make_layer(f * 4, f * 8, num_blocks[3])
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.output_conv = nn.Conv2d(f * 8, out_channels, kernel_size=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    @staticmethod
    def _make_layer(in_channels, out_channels, blocks):
        downsample = None
        if in_channels != out_channels:
            downsample = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=False),
                                       nn.BatchNorm2d(out_channels))
        layers = [ResidualBlock(in_channels, out_channels, 1, downsample)]
        for _ in range(1, blocks):
            layers.append(ResidualBlock(out_channels, out_channels))
        return nn.Sequential(*layers)

    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            eng = DeepCNNEngine(self)
            self.__dict__["_engine"] = eng
        return eng

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_engine", None)
        return state

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.B200SRError("b200sr.DeepCNN runs on CUDA sm_100a only; there is no CPU/torch fallback")
        engine = self._get_engine()
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _DeepCNNFunction.apply(self, x, *self.parameters())
        return engine.forward(x, training=self.training)


class _BN:
    def __init__(self, name, mod, C):
        self.name, self.mod, self.C = name, mod, C


class _Block:
    def __init__(self, name, mod, cin, cout):
        self.name, self.mod, self.cin, self.cout = name, mod, cin, cout
        self.ds = mod.downsample is not None


class DeepCNNEngine:
    def __init__(self, model):
        if model.in_channels != 2 or model.out_channels != 1 or model.base_features != 64 or \
                model.num_blocks != [2, 2, 2, 2]:
            raise NotImplementedError("b200sr DeepCNN engine implements the reference configuration "
                                      "DeepCNN(2, 1, [2,2,2,2], 64)")
        self.model = model
        self.blocks = []
        cin = 64
        for li, layer in enumerate((model.layer1, model.layer2, model.layer3, model.layer4)):
            cout = 64 << li
            for bi, blk in enumerate(layer):
                self.blocks.append(_Block(f"layer{li + 1}.{bi}", blk, cin, cout))
                cin = cout
        self.bns = [_BN("bn1", model.bn1, 64)]
        for b in self.blocks:
            self.bns += [_BN(b.name + ".bn1", b.mod.bn1, b.cout), _BN(b.name + ".bn2", b.mod.bn2, b.cout)]
            if b.ds:
                self.bns.append(_BN(b.name + ".downsample.1", b.mod.downsample[1], b.cout))
        self.device = None
        self.flat_p = None
        self._plans = {}
        self._eval_version = None
        self._saved = None

    # ---- flat parameter storage (same contract as UNetEngine, used by FlatAdam) -----------------------------------
    def _params(self):
        return list(self.model.parameters())

    def mark_weights_dirty(self):
        self._eval_version = None

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
                raise _lib.B200SRError(f"all DeepCNN parameters must be fp32 on {device}")
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        self.p_off, self.p_total = offs, total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_G = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad_views = []
        for p, off in zip(params, offs):
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grad_views.append(self.flat_g[off:off + p.numel()].view(p.shape))
        self.off_of = {id(p): off for p, off in zip(params, offs)}

        # packed bf16 operands (forward + dgrad) of every tensor-core conv; unpack jobs for the wgrad workspace
        convs = []  # (weight, kind_fwd, kind_dgrad, kind_unpack, cout, cin, taps)
        for b in self.blocks:
            convs.append((b.mod.conv1.weight, PACK_CONV_FWD, PACK_CONV_DGRAD, UNPACK_CONV_WGRAD, b.cout, b.cin))
            convs.append((b.mod.conv2.weight, PACK_CONV_FWD, PACK_CONV_DGRAD, UNPACK_CONV_WGRAD, b.cout, b.cout))
            if b.ds:
                convs.append((b.mod.downsample[0].weight, PACK_1X1_FWD, PACK_1X1_DGRAD, UNPACK_1X1_WGRAD, b.cout, b.cin))
        wp_total = 0
        self.wp_fwd, self.wp_dgrad = {}, {}
        for w, *_ in convs:
            self.wp_fwd[id(w)] = wp_total
            wp_total += _align(w.numel())
            self.wp_dgrad[id(w)] = wp_total
            wp_total += _align(w.numel())
        self.flat_wp = torch.zeros(wp_total, dtype=torch.bfloat16, device=device)
        pack = np.zeros(2 * len(convs), dtype=_PACK_JOB_DTYPE)
        unpack = np.zeros(len(convs), dtype=_PACK_JOB_DTYPE)
        wpb, gb, Gb = self.flat_wp.data_ptr(), self.flat_g.data_ptr(), self.flat_G.data_ptr()
        for i, (w, kf, kd, ku, cout, cin) in enumerate(convs):
            n = w.numel()
            pack[2 * i] = (w.data_ptr(), wpb + 2 * self.wp_fwd[id(w)], kf, cout, cin, 0, n)
            pack[2 * i + 1] = (w.data_ptr(), wpb + 2 * self.wp_dgrad[id(w)], kd, cout, cin, 0, n)
            off = self.off_of[id(w)]
            unpack[i] = (Gb + 4 * off, gb + 4 * off, ku, cout, cin, 0, n)
        self.pack_jobs, self.n_pack = _jobs_to_device(pack, device), len(pack)
        self.unpack_jobs, self.n_unpack = _jobs_to_device(unpack, device), len(unpack)

        ws_total, st_total = 0, 0
        self.ws_off, self.st_off = {}, {}
        for bn in self.bns:
            self.ws_off[bn.name] = ws_total
            ws_total += 4 * bn.C
            self.st_off[bn.name] = st_total
            st_total += STATS_REPLICAS * 2 * bn.C
        self.bn_ws = torch.zeros(ws_total, dtype=torch.float32, device=device)
        self.bn_stats = torch.zeros(st_total, dtype=torch.float32, device=device)
        self.bn_sums = torch.zeros(st_total, dtype=torch.float32, device=device)
        fold = np.zeros(len(self.bns), dtype=_FOLD_JOB_DTYPE)
        for i, bn in enumerate(self.bns):
            o = self.ws_off[bn.name]
            m = bn.mod
            fold[i] = (m.weight.data_ptr(), m.bias.data_ptr(), m.running_mean.data_ptr(), m.running_var.data_ptr(), 0,
                       self.bn_ws.data_ptr() + 4 * o, self.bn_ws.data_ptr() + 4 * (o + bn.C), bn.C, 0)
        self.fold_jobs = _jobs_to_device(fold, device)
        self._plans = {}
        self._eval_version = None

    def _ws(self, bn, which):
        idx = ("scale", "shift", "mean", "invstd").index(which)
        return self.bn_ws.data_ptr() + 4 * (self.ws_off[bn.name] + idx * bn.C)

    def _wp(self, table, w):
        return self.flat_wp.data_ptr() + 2 * table[id(w)]

    def _G(self, w):
        return self.flat_G.data_ptr() + 4 * self.off_of[id(w)]

    def _g(self, p):
        return self.flat_g.data_ptr() + 4 * self.off_of[id(p)]

    def _bn_of(self, name):
        return next(b for b in self.bns if b.name == name)

    def _plan(self, B, H, W):
        key = (B, H, W)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        if H % 16 != 0 or W % 16 != 0:
            raise _lib.B200SRError(f"b200sr DeepCNN needs H % 16 == 0 and W % 16 == 0 (got {H}x{W})")
        dev, bf = self.device, torch.bfloat16

        def buf(c):
            return torch.empty((B, H, W, c), dtype=bf, device=dev)

        plan = {"B": B, "H": H, "W": W, "z0": buf(64), "a0": buf(64), "p0": buf(64), "dz0": buf(64)}
        for b in self.blocks:
            for k in ("z1", "a1", "z2", "out", "dz1", "dz2"):
                plan[f"{b.name}.{k}"] = buf(b.cout)
            if b.ds:
                plan[f"{b.name}.zd"] = buf(b.cout)
                plan[f"{b.name}.dzd"] = buf(b.cout)
        plan["scratch"] = [torch.empty(B * H * W * 512, dtype=bf, device=dev) for _ in range(4)]
        self._plans[key] = plan
        return plan

    # ---- forward ----------------------------------------------------------------------------------------------------
    def _state_version(self):
        return sum(t._version for t in list(self.model.parameters()) + list(self.model.buffers()))

    def forward(self, x, training):
        if x.dim() != 4 or x.shape[1] != 2:
            raise _lib.B200SRError(f"expected input (B,2,H,W), got {tuple(x.shape)}")
        x = x.contiguous().float()
        self.ensure_ready(x.device)
        B, _, H, W = x.shape
        plan = self._plan(B, H, W)
        st = _lib.current_stream_ptr()
        npix = B * H * W
        m = self.model
        if training:
            call("b200sr_pack_jobs", self.pack_jobs.data_ptr(), self.n_pack, st)
            self.bn_stats.zero_()
        else:
            ver = self._state_version()
            if ver != self._eval_version:
                call("b200sr_pack_jobs", self.pack_jobs.data_ptr(), self.n_pack, st)
                call("b200sr_bn_fold_eval", self.fold_jobs.data_ptr(), len(self.bns), BN_EPS, st)
                self._eval_version = ver

        def stats_of(bn):
            return self.bn_stats.data_ptr() + 4 * self.st_off[bn.name] if training else None

        def finalize(bn):
            if not training:
                return
            mod = bn.mod
            track = mod.track_running_stats and mod.running_mean is not None
            call("b200sr_bn_finalize", stats_of(bn), STATS_REPLICAS, bn.C, float(npix), ptr(mod.weight), ptr(mod.bias),
                 None, BN_EPS, BN_MOMENTUM, self._ws(bn, "scale"), self._ws(bn, "shift"), self._ws(bn, "mean"),
                 self._ws(bn, "invstd"), ptr(mod.running_mean) if track else None,
                 ptr(mod.running_var) if track else None, None, st)

        def conv3(w, src, cin, cout, dst, bn):
            call("b200sr_conv3x3_fwd", ptr(src), cin, 0, cin, self._wp(self.wp_fwd, w), cout, B, H, W, ptr(dst), cout, 0,
                 None, None, 0, stats_of(bn), STATS_REPLICAS if training else 0, st)

        bn0 = self.bns[0]
        call("b200sr_conv7_fwd", ptr(x), ptr(m.conv1.weight), ptr(plan["z0"]), stats_of(bn0),
             STATS_REPLICAS if training else 0, B, H, W, st)
        finalize(bn0)
        call("b200sr_bnrelu_apply", ptr(plan["z0"]), 64, self._ws(bn0, "scale"), self._ws(bn0, "shift"), ptr(plan["a0"]),
             64, 0, None, B, H, W, st)
        call("b200sr_maxpool3x3_fwd", ptr(plan["a0"]), ptr(plan["p0"]), 64, B, H, W, st)
        xin = plan["p0"]
        for b in self.blocks:
            bn1, bn2 = self._bn_of(b.name + ".bn1"), self._bn_of(b.name + ".bn2")
            z1, a1, z2, out = (plan[f"{b.name}.{k}"] for k in ("z1", "a1", "z2", "out"))
            conv3(b.mod.conv1.weight, xin, b.cin, b.cout, z1, bn1)
            finalize(bn1)
            call("b200sr_bnrelu_apply", ptr(z1), b.cout, self._ws(bn1, "scale"), self._ws(bn1, "shift"), ptr(a1), b.cout,
                 0, None, B, H, W, st)
            conv3(b.mod.conv2.weight, a1, b.cout, b.cout, z2, bn2)
            finalize(bn2)
            if b.ds:
                bnd = self._bn_of(b.name + ".downsample.1")
                zd = plan[f"{b.name}.zd"]
                wd = b.mod.downsample[0].weight
                call("b200sr_conv1x1", ptr(xin), b.cin, 0, b.cin, self._wp(self.wp_fwd, wd), b.cout, B, H, W, ptr(zd),
                     b.cout, 0, stats_of(bnd), STATS_REPLICAS if training else 0, st)
                finalize(bnd)
                call("b200sr_bn_add_relu", ptr(z2), self._ws(bn2, "scale"), self._ws(bn2, "shift"), ptr(zd),
                     self._ws(bnd, "scale"), self._ws(bnd, "shift"), ptr(out), b.cout, npix, st)
            else:
                call("b200sr_bn_add_relu", ptr(z2), self._ws(bn2, "scale"), self._ws(bn2, "shift"), ptr(xin), None, None,
                     ptr(out), b.cout, npix, st)
            xin = out
        y = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        oc = m.output_conv
        call("b200sr_headw_fwd", ptr(xin), 512, ptr(oc.weight), ptr(oc.bias), ptr(y), npix, st)
        if training:
            bufs = [bn.mod.num_batches_tracked for bn in self.bns if bn.mod.num_batches_tracked is not None]
            if bufs:
                torch._foreach_add_(bufs, 1)
            self._saved = (plan, x)
        return y

    # ---- backward ---------------------------------------------------------------------------------------------------
    def backward(self, dout, bucket_hook=None, want_dx=False):
        if self._saved is None:
            raise _lib.B200SRError("backward() without a preceding train-mode forward")
        if want_dx:
            raise NotImplementedError("DeepCNN input gradients are not needed by the reference training step")
        plan, x = self._saved
        B, H, W = plan["B"], plan["H"], plan["W"]
        npix = B * H * W
        st = _lib.current_stream_ptr()
        m = self.model
        dout = dout.contiguous().float()
        self.flat_g.zero_()
        self.flat_G.zero_()
        self.bn_sums.zero_()
        sc = [t.data_ptr() for t in plan["scratch"]]
        oc = m.output_conv
        last = plan[f"{self.blocks[-1].name}.out"]
        call("b200sr_headw_bwd", ptr(dout), ptr(last), 512, ptr(oc.weight), sc[0], self._g(oc.weight), self._g(oc.bias),
             npix, st)
        d_out = sc[0]

        def sums_of(bn):
            return self.bn_sums.data_ptr() + 4 * self.st_off[bn.name]

        def bn_bwd_relu(bn, dy, z, dz):
            """standard BN+ReLU backward (mask from scale*z+shift > 0)"""
            args = (self._ws(bn, "scale"), self._ws(bn, "shift"), self._ws(bn, "mean"), self._ws(bn, "invstd"))
            call("b200sr_bn_bwd_reduce", dy, bn.C, 0, ptr(z), bn.C, *args, sums_of(bn), STATS_REPLICAS, npix, st)
            call("b200sr_bn_bwd_apply_fused", dy, bn.C, 0, ptr(z), bn.C, *args, sums_of(bn), STATS_REPLICAS, float(npix),
                 self._g(bn.mod.weight), self._g(bn.mod.bias), ptr(dz), npix, st)

        def bn_bwd_masked(bn, dy, z, mask, dz):
            call("b200sr_bn_bwd_masked", dy, ptr(z), ptr(mask), bn.C, self._ws(bn, "scale"), self._ws(bn, "shift"),
                 self._ws(bn, "mean"), self._ws(bn, "invstd"), sums_of(bn), STATS_REPLICAS, float(npix),
                 self._g(bn.mod.weight), self._g(bn.mod.bias), ptr(dz), npix, st)

        for bi in range(len(self.blocks) - 1, -1, -1):
            b = self.blocks[bi]
            xin = plan["p0"] if bi == 0 else plan[f"{self.blocks[bi - 1].name}.out"]
            bn1, bn2 = self._bn_of(b.name + ".bn1"), self._bn_of(b.name + ".bn2")
            z1, a1, z2, out, dz1, dz2 = (plan[f"{b.name}.{k}"] for k in ("z1", "a1", "z2", "out", "dz1", "dz2"))
            free = [p for p in sc if p != d_out]
            d_a1, dx, dxd = free[0], free[1], free[2]
            w1, w2 = b.mod.conv1.weight, b.mod.conv2.weight
            bn_bwd_masked(bn2, d_out, z2, out, dz2)
            call("b200sr_conv3x3_wgrad", ptr(a1), b.cout, 0, b.cout, ptr(dz2), b.cout, 0, b.cout, B, H, W, self._G(w2), st)
            call("b200sr_conv3x3_dgrad", ptr(dz2), b.cout, 0, b.cout, self._wp(self.wp_dgrad, w2), b.cout, B, H, W, d_a1,
                 b.cout, 0, None, 0, st)
            bn_bwd_relu(bn1, d_a1, z1, dz1)
            call("b200sr_conv3x3_wgrad", ptr(xin), b.cin, 0, b.cin, ptr(dz1), b.cout, 0, b.cout, B, H, W, self._G(w1), st)
            call("b200sr_conv3x3_dgrad", ptr(dz1), b.cout, 0, b.cout, self._wp(self.wp_dgrad, w1), b.cin, B, H, W, dx,
                 b.cin, 0, None, 0, st)
            d_in = d_a1  # dead by now
            if b.ds:
                bnd = self._bn_of(b.name + ".downsample.1")
                zd, dzd = plan[f"{b.name}.zd"], plan[f"{b.name}.dzd"]
                wd = b.mod.downsample[0].weight
                bn_bwd_masked(bnd, d_out, zd, out, dzd)
                call("b200sr_conv1x1_wgrad", ptr(xin), b.cin, 0, b.cin, ptr(dzd), b.cout, 0, b.cout, B, H, W, self._G(wd),
                     st)
                call("b200sr_conv1x1", ptr(dzd), b.cout, 0, b.cout, self._wp(self.wp_dgrad, wd), b.cin, B, H, W, dxd,
                     b.cin, 0, None, 0, st)
                call("b200sr_add_masked", dx, dxd, None, d_in, npix * b.cin, st)
            else:
                call("b200sr_add_masked", dx, d_out, ptr(out), d_in, npix * b.cin, st)
            d_out = d_in
        # stem
        free = [p for p in sc if p != d_out]
        d_a0 = free[0]
        call("b200sr_maxpool3x3_bwd", ptr(plan["a0"]), d_out, d_a0, 64, B, H, W, st)
        bn_bwd_relu(self.bns[0], d_a0, plan["z0"], plan["dz0"])
        call("b200sr_conv7_wgrad", ptr(x), ptr(plan["dz0"]), self._g(m.conv1.weight), B, H, W, st)
        call("b200sr_pack_jobs", self.unpack_jobs.data_ptr(), self.n_unpack, st)
        if bucket_hook is not None:
            bucket_hook(0, self.p_total)
        self._saved = None
        return self.grad_views


class DeepCNNTrainer:
    """MSE + Adam(lr 1e-4) train step of the DeepCNN baseline (results/deepcnn_history.json `config`)."""

    def __init__(self, model, device="cuda", learning_rate=1e-4, model_save_dir="models", verbose=True):
        from .losses import CombinedLoss
        from .optim import FlatAdam
        self.model = model.to(device)
        self.device = device
        self.criterion = CombinedLoss(mse_weight=1.0, ssim_weight=0.0)
        self.optimizer = FlatAdam(self.model, lr=learning_rate)
        self.model_save_dir = Path(model_save_dir)
        self.model_save_dir.mkdir(parents=True, exist_ok=True)
        self._reducer = None
        if verbose:
            print(f"Total parameters: {sum(p.numel() for p in self.model.parameters()):,}")

    def train_step(self, inputs, targets):
        self.model.train()
        engine = self.model._get_engine()
        self.optimizer.host_pre_step()
        out = engine.forward(inputs, training=True)
        loss, dout = self.criterion.value_and_grad(out, targets)
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
        self.optimizer.device_step(grad_scale=scale)
        return loss

    def save_checkpoint(self, epoch, val_loss, is_best=False):
        ck = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "val_loss": val_loss}
        if is_best:
            torch.save(ck, self.model_save_dir / "deepcnn_best.pt")
        torch.save(ck, self.model_save_dir / "deepcnn_latest.pt")
