"""Execution engine of the UNet hot path: owns the HBM layout and issues the C-ABI ops in order.

Reference being replaced: UNet.forward + autograd backward (/root/reference/src/unet_model.py:82-118).

HBM layout (all activations NHWC bf16, resident for the whole step):
  * parameters: one flat fp32 buffer; every nn.Parameter is a view into it (state_dict layout unchanged);
    gradients: one flat fp32 buffer with the same offsets (bucketable for the NCCL all-reduce).
  * derived bf16 operand copies of the conv weights (forward and dgrad packings) in one flat buffer,
    refreshed by ONE table-driven kernel launch; never part of the state_dict.
  * per decoder level one concat buffer (B,H,W,2C): ConvTranspose writes channels [0,C), the encoder block's
    BN+ReLU pass writes channels [C,2C) -> torch.cat never runs.
  * per conv: raw output z (kept for backward), per BN: scale/shift/mean/invstd and statistics replicas.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

STATS_REPLICAS = 4   # legacy (atomic) statistic replicas, still used by the DeepCNN / Fast-DDPM engines
WGRAD_WS_FLOATS = 24 * 1024 * 1024  # 96 MB of split-K partials (largest layer: 57 MB); caps the split factor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

PACK_CONV_FWD, PACK_CONV_DGRAD, PACK_CONVT_FWD, PACK_CONVT_DGRAD, UNPACK_CONV_WGRAD, UNPACK_CONVT_WGRAD = range(6)
PACK_CONV_BOTH, PACK_CONVT_BOTH = 9, 10  # forward + dgrad packing from one read of the parameter
PACK_CONV_FWD_SPLIT3, PACK_CONVT_FWD_SPLIT3 = 11, 12  # bf16x3 operand split of the fp32-accuracy eval mode

_PACK_JOB_DTYPE = np.dtype([("src", "<u8"), ("dst", "<u8"), ("kind", "<i4"), ("cout", "<i4"), ("cin", "<i4"),
                            ("pad", "<i4"), ("count", "<i8")])
_FOLD_JOB_DTYPE = np.dtype([("gamma", "<u8"), ("beta", "<u8"), ("rmean", "<u8"), ("rvar", "<u8"), ("cbias", "<u8"),
                            ("scale", "<u8"), ("shift", "<u8"), ("C", "<i4"), ("pad", "<i4")])


class _Nvtx:
    """NVTX ranges around every layer of the forward / backward pass (B200SR_NVTX=1): shows up in nsys / ncu --nvtx as
    fwd/<layer>, bwd/<block>, so a profile maps kernels to reference layers. Off by default (two host calls per range)."""

    enabled = __import__("os").environ.get("B200SR_NVTX") is not None

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Nvtx.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


def _jobs_to_device(arr: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(arr.view(np.uint8).copy()).to(device)


class ConvSpec:
    """One Conv3x3+BN+ReLU layer of a UNetBlock (unet_model.py:27-32)."""

    def __init__(self, name, conv, bn, cin, cout, level):
        self.name, self.conv, self.bn, self.cin, self.cout, self.level = name, conv, bn, cin, cout, level


class UpSpec:
    """One ConvTranspose2d(k2,s2) (unet_model.py:67-76)."""

    def __init__(self, name, mod, cin, cout, level):
        self.name, self.mod, self.cin, self.cout, self.level = name, mod, cin, cout, level


class UNetEngine:
    def __init__(self, model):
        self.model = model
        f = model.init_features
        if model.in_channels != 2 or model.out_channels != 1 or f != 64:
            raise NotImplementedError(
                "b200sr UNet engine implements the reference configuration UNet(in_channels=2, out_channels=1, "
                f"init_features=64); got ({model.in_channels}, {model.out_channels}, {f})")
        m = model
        # the 1x1 head is `final_conv` in UNet (unet_model.py:80) and `final` in UNetStage (ModelLoader.py:186)
        self.head = m.final_conv if hasattr(m, "final_conv") else m.final
        chans = [f, 2 * f, 4 * f, 8 * f, 16 * f]
        self.chans = chans
        blocks = [("enc1", m.enc1, 2, chans[0], 0), ("enc2", m.enc2, chans[0], chans[1], 1),
                  ("enc3", m.enc3, chans[1], chans[2], 2), ("enc4", m.enc4, chans[2], chans[3], 3),
                  ("bottleneck", m.bottleneck, chans[3], chans[4], 4),
                  ("dec4", m.dec4, chans[4], chans[3], 3), ("dec3", m.dec3, chans[3], chans[2], 2),
                  ("dec2", m.dec2, chans[2], chans[1], 1), ("dec1", m.dec1, chans[1], chans[0], 0)]
        self.blocks = {}
        self.convs = []  # forward order
        for name, blk, cin, cout, level in blocks:
            c1 = ConvSpec(f"{name}.conv.0", blk.conv[0], blk.conv[1], cin, cout, level)
            c2 = ConvSpec(f"{name}.conv.3", blk.conv[3], blk.conv[4], cout, cout, level)
            self.blocks[name] = (c1, c2)
            self.convs += [c1, c2]
        self.ups = {4: UpSpec("upconv4", m.upconv4, chans[4], chans[3], 3),
                    3: UpSpec("upconv3", m.upconv3, chans[3], chans[2], 2),
                    2: UpSpec("upconv2", m.upconv2, chans[2], chans[1], 1),
                    1: UpSpec("upconv1", m.upconv1, chans[1], chans[0], 0)}
        self.device = None
        self.flat_p = None
        self._plans = {}
        self._eval_version = None
        self._saved = None  # (plan, x) of the last train-mode forward
        self._side = None   # side stream for weight gradients (created on first backward)
        self._events = []
        import os
        self.overlap_wgrad = os.environ.get("B200SR_NO_OVERLAP") is None
        self.hp_chain = os.environ.get("B200SR_HP") is not None
        # experiment switch (off): order each layer's wgrad behind its dgrad so it overlaps the NEXT BatchNorm backward;
        # measured 13.20 vs 12.78 ms per step — the following dgrad then waits for the wgrad CTAs to retire
        self.wgrad_late = os.environ.get("B200SR_WGRAD_LATE") is not None
        self.timing_skip_pack = os.environ.get("B200SR_TIMING_SKIP_PACK") is not None  # timing experiment only
        self.eval_graphs = os.environ.get("B200SR_NO_EVAL_GRAPH") is None
        self.fused_bn_finalize = os.environ.get("B200SR_NO_FUSED_BN") is None  # A/B switch: separate b200sr_bn_finalize launches
        self.fused_pool_bnred = os.environ.get("B200SR_NO_FUSED_POOL_BNRED") is None  # A/B switch: separate reduction pass
        # head backward fused with the BatchNorm-backward reduction of dec1.conv.3 (-0.1 ms per step once the kernel ran
        # without register spills); A/B switch: B200SR_NO_FUSED_HEAD_BNRED=1 -> the two separate kernels
        self.fused_head_bnred = os.environ.get("B200SR_NO_FUSED_HEAD_BNRED") is None
        self._ab_full_stats = os.environ.get("B200SR_DGRAD_FULL_STATS") is not None  # A/B switch: sums + squares of all columns
        self._eval_graph_cache, self._eval_graph_calls = {}, {}
        self._hp = None

    # ------------------------------------------------------------------------------------------------
    # parameter flattening and derived operand buffers
    # ------------------------------------------------------------------------------------------------
    def _params(self):
        ps = self.__dict__.get("_param_list")
        if ps is None:
            ps = self._param_list = list(self.model.parameters())
        return ps

    def _is_flat(self) -> bool:
        """True while every parameter still is the view of the flat buffer handed out by ensure_ready() and the BatchNorm
        buffers are where the job tables expect them. A complete scan: 66 data_ptr() reads per call (~20 us), so a
        replaced p.data anywhere in the model is noticed at the next call, not up to 64 calls later."""
        if self.flat_p is None:
            return False
        params = list(self.model.parameters())
        if len(params) != len(self.p_off) or params[0].device != self.flat_p.device:
            return False
        if tuple(p.data_ptr() for p in params) != self._flat_ptrs:
            return False
        self._param_list = params
        # raw buffer pointers are baked into the fold-job table
        return self._fold_ptrs == [(cs.bn.running_mean.data_ptr(), cs.bn.running_var.data_ptr()) for cs in self.convs]

    def ensure_ready(self, device):
        """(Re)build flat parameter storage and derived buffers if parameters moved (model.to(), new tensors)."""
        if self.device == device and self._is_flat():
            return
        _lib.require_device()
        self.device = device
        self.__dict__.pop("_param_list", None)
        self.__dict__.pop("_pindex", None)
        params = self._params()
        for p in params:
            if p.device != device or p.dtype != torch.float32:
                raise _lib.B200SRError(f"all UNet parameters must be fp32 on {device} (got {p.dtype} on {p.device})")
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        self.p_off, self.p_total = offs, total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=device)
        from .ddp import alloc_comm_buffer
        self.flat_g = alloc_comm_buffer(total, device)  # NCCL-registered when data-parallel (zero-copy / NVLS all-reduce)
        self.flat_g_registered = self.flat_g is not None
        if self.flat_g is None:
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=device)
        # (flat_g is zeroed ONCE: every real gradient is WRITTEN by a deterministic kernel each step; the conv biases in front
        # of a BatchNorm have an exactly-zero gradient and are never touched)
        self.grad_views = []
        for p, off in zip(params, offs):
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grad_views.append(self.flat_g[off:off + p.numel()].view(p.shape))
        self.off_of = {id(p): off for p, off in zip(params, offs)}
        self._flat_ptrs = tuple(p.data_ptr() for p in params)

        # derived bf16 operand copies: forward + dgrad packings of every tensor-core layer
        wp_total = 0
        self.wp_fwd, self.wp_dgrad = {}, {}
        for cs in self.convs[1:]:
            n = cs.cout * cs.cin * 9
            self.wp_fwd[cs.name] = wp_total
            wp_total += _align(n)
            self.wp_dgrad[cs.name] = wp_total
            wp_total += _align(n)
        for us in self.ups.values():
            n = us.cin * us.cout * 4
            self.wp_fwd[us.name] = wp_total
            wp_total += _align(n)
            self.wp_dgrad[us.name] = wp_total
            wp_total += _align(n)
        self.flat_wp = torch.zeros(wp_total, dtype=torch.bfloat16, device=device)

        # one job per layer writes BOTH packings (forward + dgrad) from a single read of the fp32 parameter
        pack = np.zeros(len(self.convs) - 1 + len(self.ups), dtype=_PACK_JOB_DTYPE)
        i = 0
        wp_base = self.flat_wp.data_ptr()
        for cs in self.convs[1:]:
            w = cs.conv.weight
            pack[i] = (w.data_ptr(), wp_base + 2 * self.wp_fwd[cs.name], PACK_CONV_BOTH, cs.cout, cs.cin, 0,
                       wp_base + 2 * self.wp_dgrad[cs.name])
            i += 1
        for us in self.ups.values():
            w = us.mod.weight
            pack[i] = (w.data_ptr(), wp_base + 2 * self.wp_fwd[us.name], PACK_CONVT_BOTH, us.cout, us.cin, 0,
                       wp_base + 2 * self.wp_dgrad[us.name])
            i += 1
        # two contiguous groups so the train step can stage the packing: enc1/enc2 (needed first, tiny) on the main
        # stream, everything else on the side stream while those layers run
        early_dst = {wp_base + 2 * self.wp_fwd[cs.name] for cs in self.convs[1:4]}
        early = np.array([int(d) in early_dst for d in pack["dst"]])
        order = np.concatenate([np.nonzero(early)[0], np.nonzero(~early)[0]])
        pack = pack[order]
        self.pack_groups = (int(early.sum()), int((~early).sum()), 0)
        self.pack_jobs = _jobs_to_device(pack, device)
        self.n_pack = len(pack)

        # per-BN workspace: scale, shift, mean, invstd, eval scale, eval shift; forward statistic slots (one per CTA:
        # deterministic, see include/b200sr.h) and the final backward sums [2][C]
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        self.n_sms = sms
        ws_total, st_total, sm_total = 0, 0, 0
        self.bn_ws_off, self.bn_st_off, self.bn_sum_off, self.bn_slots = {}, {}, {}, {}
        for li, cs in enumerate(self.convs):
            self.bn_ws_off[cs.name] = ws_total
            ws_total += 6 * cs.cout
            self.bn_slots[cs.name] = 2 * sms if li == 0 else sms  # first conv: 2 CTAs per SM, persistent convs: 1
            self.bn_st_off[cs.name] = st_total
            st_total += self.bn_slots[cs.name] * 2 * cs.cout
            self.bn_sum_off[cs.name] = sm_total
            sm_total += 2 * cs.cout
        self.bn_ws = torch.zeros(ws_total, dtype=torch.float32, device=device)
        self.bn_stats = torch.zeros(st_total, dtype=torch.float32, device=device)
        self.bn_sums = torch.zeros(sm_total, dtype=torch.float32, device=device)
        # column-sum slots of the decoder concat gradients (ConvTranspose bias gradient), one slot per dgrad CTA
        self.up_st_off, up_total = {}, 0
        for k, us in self.ups.items():
            self.up_st_off[k] = up_total
            up_total += sms * 2 * 2 * us.cout
        self.up_stats = torch.zeros(up_total, dtype=torch.float32, device=device)
        # workspaces of the deterministic reductions. Main-stream ops (BatchNorm backward, head) and side-stream ops
        # (weight gradients) never share one; ticket counters are zeroed once and reset by the kernels themselves.
        bn_ws_floats = max(int(call("b200sr_bn_bwd_ws_floats", cs.cout)) for cs in self.convs)
        self.red_ws = torch.empty(max(bn_ws_floats, 4 * sms * 72, 3 * sms * 200), dtype=torch.float32, device=device)
        self.red_counters = torch.zeros(64, dtype=torch.int32, device=device)
        self.bn_fwd_counters = torch.zeros(16 * len(self.convs), dtype=torch.int32, device=device)  # fused finalize tickets
        self.wg_ws = torch.empty(WGRAD_WS_FLOATS, dtype=torch.float32, device=device)

        fold = np.zeros(len(self.convs), dtype=_FOLD_JOB_DTYPE)
        for i, cs in enumerate(self.convs):
            o = self.bn_ws_off[cs.name]
            fold[i] = (cs.bn.weight.data_ptr(), cs.bn.bias.data_ptr(), cs.bn.running_mean.data_ptr(),
                       cs.bn.running_var.data_ptr(), cs.conv.bias.data_ptr() if cs.conv.bias is not None else 0,
                       self.bn_ws.data_ptr() + 4 * (o + 4 * cs.cout), self.bn_ws.data_ptr() + 4 * (o + 5 * cs.cout),
                       cs.cout, 0)
        self._fold_np = fold
        self.fold_jobs = _jobs_to_device(fold, device)
        self._fold_ptrs = [(cs.bn.running_mean.data_ptr(), cs.bn.running_var.data_ptr()) for cs in self.convs]
        self._plans = {}
        self._eval_graph_cache, self._eval_graph_calls = {}, {}  # captured graphs point into the old buffers
        self._eval_version = None

    def _bn(self, cs, which):
        """Pointer (int) into the BN workspace: train-mode scale, shift, mean, invstd (kept for backward) and the
        eval-mode folded scale / shift, which live in their own slots so that an eval forward between a train forward
        and its backward leaves the saved train-mode values alone."""
        idx = ("scale", "shift", "mean", "invstd", "escale", "eshift").index(which)
        return self.bn_ws.data_ptr() + 4 * (self.bn_ws_off[cs.name] + idx * cs.cout)

    def _wp(self, table, name):
        return self.flat_wp.data_ptr() + 2 * table[name]

    def _state_version(self):
        v = 0
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            v += t._version
        return v

    def mark_weights_dirty(self):
        """Parameters were changed behind torch's back (raw-pointer kernels): drop eval-mode derived state."""
        self._eval_version = None

    def repack_weights(self):
        call("b200sr_pack_jobs", self.pack_jobs.data_ptr(), self.n_pack, _lib.current_stream_ptr())

    def _repack_staged(self):
        """Train step: pack what enc1/enc2 need on the current stream, the rest on the side stream, overlapping
        the first layers. Returns the event the forward must wait on before enc3; the dgrad packings are joined at
        the start of backward (self._pack_done)."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        n0, n1, n2 = self.pack_groups
        isz = _PACK_JOB_DTYPE.itemsize
        base = self.pack_jobs.data_ptr()
        fork = torch.cuda.Event()
        fork.record(main)  # parameters are final (Adam of the previous step) at this point of the main stream
        self._side.wait_event(fork)
        call("b200sr_pack_jobs", base, n0, main.cuda_stream)
        sst = self._side.cuda_stream
        call("b200sr_pack_jobs", base + n0 * isz, n1, sst)
        fwd_done = torch.cuda.Event()
        fwd_done.record(self._side)
        if n2 > 0:
            call("b200sr_pack_jobs", base + (n0 + n1) * isz, n2, sst)
            self._pack_done = torch.cuda.Event()
            self._pack_done.record(self._side)
        else:
            self._pack_done = fwd_done  # the dgrad packings came with the forward ones
        return fwd_done

    # ------------------------------------------------------------------------------------------------
    # activation plans
    # ------------------------------------------------------------------------------------------------
    def _plan(self, B, H, W, train):
        key = (B, H, W, train)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        if H % 16 != 0 or W % 16 != 0 or H <= 0 or W <= 0:
            # four 2x2 poolings, like the reference (src/unet_model.py:56-75: any multiple of 16). Levels whose size is not a
            # multiple of the GEMM pixel tile run the same kernels with ragged edge tiles (csrc/conv3x3.cuh).
            raise _lib.B200SRError(f"b200sr UNet needs H % 16 == 0 and W % 16 == 0 (got {H}x{W})")
        dev, bf = self.device, torch.bfloat16
        ch = self.chans
        plan = {"B": B, "H": H, "W": W}

        def buf(h, w, c):
            return torch.empty((B, h, w, c), dtype=bf, device=dev)

        for lvl in range(5):
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            if lvl < 4:
                plan[f"cat{lvl}"] = buf(h, w, 2 * c)   # [0,C) upconv output, [C,2C) encoder skip
                plan[f"pool{lvl}"] = buf(h // 2, w // 2, c)
                plan[f"enc_a1_{lvl}"] = buf(h, w, c)
                plan[f"dec_a1_{lvl}"] = buf(h, w, c)
                plan[f"dec_a2_{lvl}"] = buf(h, w, c)
            else:
                plan["bot_a1"] = buf(h, w, c)
                plan["bot_a2"] = buf(h, w, c)
        if train:
            for cs in self.convs:
                h, w = H >> cs.level, W >> cs.level
                plan["z:" + cs.name] = buf(h, w, cs.cout)
                plan["dz:" + cs.name] = buf(h, w, cs.cout)  # own buffer per layer: read by the side-stream wgrad
            for lvl in range(4):
                h, w, c = H >> lvl, W >> lvl, ch[lvl]
                plan[f"dcat{lvl}"] = buf(h, w, 2 * c)
            big = B * H * W * ch[0]
            plan["scratch"] = [torch.empty(big, dtype=bf, device=dev) for _ in range(3)]
        self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------------
    def _check_input(self, x):
        if x.dim() != 4 or x.shape[1] != 2:
            raise _lib.B200SRError(f"expected input (B,2,H,W), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise _lib.B200SRError("b200sr UNet runs on CUDA sm_100a only; there is no CPU path")
        return x.contiguous().float()

    def forward_eval(self, x, precision="bf16"):
        """Eval-mode forward. precision 'bf16' (default): bf16 operands / activations, fp32 accumulation. precision
        'fp32': the reference's inference arithmetic (an fp32 forward, VolumeVisualization.py:932-964) to rel-L2 <= 1e-4 on
        the same tensor-core kernels through bf16x3 operand splitting (csrc/conv3x3.cuh SPLIT; 3x the tensor work).
        From the third call with a given input shape the ~30 launches are replayed from a CUDA graph (small inference
        batches are launch-latency bound between dependent kernels): the input is copied into the graph's static buffer
        and the result is returned as a fresh tensor. B200SR_NO_EVAL_GRAPH=1 disables it."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"Unknown precision: {precision}. Choose from: ['bf16', 'fp32']")
        x = self._check_input(x)
        self.ensure_ready(x.device)
        self._refresh_eval_weights(precision)
        launch = self._forward_eval_launch if precision == "bf16" else self._forward_eval_launch_fp32
        if not self.eval_graphs or torch.cuda.is_current_stream_capturing():
            return launch(x)
        key = (tuple(x.shape), x.device, precision)
        g = self._eval_graph_cache.get(key)
        if g is None:
            n = self._eval_graph_calls.get(key, 0)
            self._eval_graph_calls[key] = n + 1
            if n < 2:  # lazy one-time set-up (plan buffers, TMA descriptors, kernel attributes) happens eagerly
                return launch(x)
            static_x = x.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = launch(static_x)
            g = (graph, static_x, static_out)
            self._eval_graph_cache[key] = g
        graph, static_x, static_out = g
        static_x.copy_(x)
        graph.replay()
        return static_out.clone()

    def _refresh_eval_weights(self, precision="bf16"):
        """Eval-mode derived state (packed bf16 weights, folded BatchNorm) follows the parameters / buffers."""
        ver = self._state_version()
        if ver != self._eval_version:
            self.repack_weights()
            call("b200sr_bn_fold_eval", self.fold_jobs.data_ptr(), len(self.convs), BN_EPS, _lib.current_stream_ptr())
            self._eval_version = ver
        if precision == "fp32" and ver != getattr(self, "_eval_version_fp32", None):
            self._pack_split_weights()
            self._eval_version_fp32 = ver

    def _pack_split_weights(self):
        """[w_hi | w_hi | w_lo] packings of every tensor-core layer for the fp32-accuracy eval mode (built on first use)."""
        if getattr(self, "_wp32_for", None) != self.flat_p.data_ptr():
            total, self.wp32 = 0, {}
            for cs in self.convs[1:]:
                self.wp32[cs.name] = total
                total += _align(3 * cs.cout * cs.cin * 9)
            for us in self.ups.values():
                self.wp32[us.name] = total
                total += _align(3 * us.cin * us.cout * 4)
            self.flat_wp32 = torch.zeros(total, dtype=torch.bfloat16, device=self.device)
            jobs = np.zeros(len(self.convs) - 1 + len(self.ups), dtype=_PACK_JOB_DTYPE)
            base = self.flat_wp32.data_ptr()
            i = 0
            for cs in self.convs[1:]:
                w = cs.conv.weight
                jobs[i] = (w.data_ptr(), base + 2 * self.wp32[cs.name], PACK_CONV_FWD_SPLIT3, cs.cout, cs.cin, 0, 3 * w.numel())
                i += 1
            for us in self.ups.values():
                w = us.mod.weight
                jobs[i] = (w.data_ptr(), base + 2 * self.wp32[us.name], PACK_CONVT_FWD_SPLIT3, us.cout, us.cin, 0,
                           3 * w.numel())
                i += 1
            self.pack32_jobs, self.n_pack32 = _jobs_to_device(jobs, self.device), len(jobs)
            self._wp32_for = self.flat_p.data_ptr()
        call("b200sr_pack_jobs", self.pack32_jobs.data_ptr(), self.n_pack32, _lib.current_stream_ptr())

    def _forward_eval_launch_fp32(self, x):
        """The eval forward with every activation stored as (B,H,W,3C) bf16 [hi | lo | hi] and every weight as
        [w_hi | w_hi | w_lo]: fp32 accuracy (2^-17 operands, fp32 accumulation, fp32 BatchNorm fold) on the bf16 tensor
        cores. Decoder concat buffers hold 2C logical channels, i.e. parts 2C apart; ConvTranspose writes [0,C) of each
        part, the encoder skip writes [C,2C)."""
        B, _, H, W = x.shape
        key = (B, H, W, "fp32")
        plan = self._plans.get(key)
        ch = self.chans
        if plan is None:
            if H % 16 != 0 or W % 16 != 0:
                raise _lib.B200SRError(f"b200sr UNet needs H % 16 == 0 and W % 16 == 0 (got {H}x{W})")
            bf, dev = torch.bfloat16, self.device
            plan = {}
            for lvl in range(5):
                h, w, c = H >> lvl, W >> lvl, ch[lvl]
                if lvl < 4:
                    plan[f"cat{lvl}"] = torch.empty((B, h, w, 6 * c), dtype=bf, device=dev)
                    plan[f"pool{lvl}"] = torch.empty((B, h // 2, w // 2, 3 * c), dtype=bf, device=dev)
                    for k in ("enc_a1", "dec_a1", "dec_a2"):
                        plan[f"{k}_{lvl}"] = torch.empty((B, h, w, 3 * c), dtype=bf, device=dev)
                else:
                    plan["bot_a1"] = torch.empty((B, h, w, 3 * c), dtype=bf, device=dev)
                    plan["bot_a2"] = torch.empty((B, h, w, 3 * c), dtype=bf, device=dev)
            self._plans[key] = plan
        st = _lib.current_stream_ptr()
        wbase = self.flat_wp32.data_ptr()

        def conv(cs, src, s_stride, cin_logical, dst, d_stride, d_off, part, h, w):
            call("b200sr_conv3x3_fwd_split", ptr(src), s_stride, 0, 3 * cin_logical, wbase + 2 * self.wp32[cs.name],
                 cs.cout, B, h, w, ptr(dst), d_stride, d_off, part, self._bn(cs, "escale"), self._bn(cs, "eshift"), 1, st)

        cur = None
        for lvl, name in enumerate(["enc1", "enc2", "enc3", "enc4"]):
            c1, c2 = self.blocks[name]
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            a1, cat, pool = plan[f"enc_a1_{lvl}"], plan[f"cat{lvl}"], plan[f"pool{lvl}"]
            if lvl == 0:
                call("b200sr_conv1_fwd_split", ptr(x), ptr(c1.conv.weight), self._bn(c1, "escale"), self._bn(c1, "eshift"),
                     1, ptr(a1), B, h, w, st)
            else:
                conv(c1, cur, 3 * c1.cin, c1.cin, a1, 3 * c, 0, c, h, w)
            conv(c2, a1, 3 * c, c, cat, 6 * c, c, 2 * c, h, w)
            call("b200sr_maxpool2x2_fwd_split", ptr(cat), 6 * c, c, 2 * c, c, ptr(pool), B, h, w, st)
            cur = pool
        c1, c2 = self.blocks["bottleneck"]
        h, w, c = H >> 4, W >> 4, ch[4]
        conv(c1, cur, 3 * c1.cin, c1.cin, plan["bot_a1"], 3 * c, 0, c, h, w)
        conv(c2, plan["bot_a1"], 3 * c, c, plan["bot_a2"], 3 * c, 0, c, h, w)
        cur = plan["bot_a2"]
        for k in (4, 3, 2, 1):
            us = self.ups[k]
            lvl = us.level
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            cat = plan[f"cat{lvl}"]
            call("b200sr_convT2x2_fwd_split", ptr(cur), 3 * us.cin, 0, 3 * us.cin, wbase + 2 * self.wp32[us.name], us.cout,
                 ptr(us.mod.bias), B, h // 2, w // 2, ptr(cat), 6 * c, 0, 2 * c, st)
            c1, c2 = self.blocks[f"dec{k}"]
            conv(c1, cat, 6 * c, 2 * c, plan[f"dec_a1_{lvl}"], 3 * c, 0, c, h, w)
            conv(c2, plan[f"dec_a1_{lvl}"], 3 * c, c, plan[f"dec_a2_{lvl}"], 3 * c, 0, c, h, w)
            cur = plan[f"dec_a2_{lvl}"]
        fc = self.head
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        call("b200sr_head_fwd_split", ptr(cur), ptr(fc.weight), ptr(fc.bias), ptr(out), B * H * W, st)
        return out

    def _forward_eval_launch(self, x):
        B, _, H, W = x.shape
        plan = self._plan(B, H, W, False)
        st = _lib.current_stream_ptr()
        ch = self.chans

        def conv(cs, src, s_stride, s_off, dst, d_stride, d_off, h, w):
            call("b200sr_conv3x3_fwd", ptr(src), s_stride, s_off, cs.cin, self._wp(self.wp_fwd, cs.name), cs.cout,
                 B, h, w, ptr(dst), d_stride, d_off, self._bn(cs, "escale"), self._bn(cs, "eshift"), 1, None, 0, st)

        cur = None
        for lvl, name in enumerate(["enc1", "enc2", "enc3", "enc4"]):
            c1, c2 = self.blocks[name]
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            a1, cat, pool = plan[f"enc_a1_{lvl}"], plan[f"cat{lvl}"], plan[f"pool{lvl}"]
            if lvl == 0:
                call("b200sr_conv1_fwd", ptr(x), ptr(c1.conv.weight), self._bn(c1, "escale"), self._bn(c1, "eshift"), 1,
                     ptr(a1), None, 0, B, h, w, st)
            else:
                conv(c1, cur, c1.cin, 0, a1, c, 0, h, w)
            conv(c2, a1, c, 0, cat, 2 * c, c, h, w)
            call("b200sr_maxpool2x2_fwd", ptr(cat), 2 * c, c, c, ptr(pool), B, h, w, st)
            cur = pool
        c1, c2 = self.blocks["bottleneck"]
        h, w, c = H >> 4, W >> 4, ch[4]
        conv(c1, cur, c1.cin, 0, plan["bot_a1"], c, 0, h, w)
        conv(c2, plan["bot_a1"], c, 0, plan["bot_a2"], c, 0, h, w)
        cur = plan["bot_a2"]
        for k in (4, 3, 2, 1):
            us = self.ups[k]
            lvl = us.level
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            cat = plan[f"cat{lvl}"]
            call("b200sr_convT2x2_fwd", ptr(cur), us.cin, 0, us.cin, self._wp(self.wp_fwd, us.name), us.cout,
                 ptr(us.mod.bias), B, h // 2, w // 2, ptr(cat), 2 * c, 0, st)
            c1, c2 = self.blocks[f"dec{k}"]
            conv(c1, cat, 2 * c, 0, plan[f"dec_a1_{lvl}"], c, 0, h, w)
            conv(c2, plan[f"dec_a1_{lvl}"], c, 0, plan[f"dec_a2_{lvl}"], c, 0, h, w)
            cur = plan[f"dec_a2_{lvl}"]
        fc = self.head
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        call("b200sr_head_fwd", ptr(cur), ptr(fc.weight), ptr(fc.bias), ptr(out), B * H * W, st)
        return out

    def _conv_bn_train(self, plan, cs, src, s_stride, s_off, h, w, act, a_stride, a_off, pooled, x_input=None):
        """conv -> batch statistics -> finalize -> BN-apply+ReLU(+pool)."""
        with _Nvtx("fwd/" + cs.name):
            self._conv_bn_train_impl(plan, cs, src, s_stride, s_off, h, w, act, a_stride, a_off, pooled, x_input)

    def _conv_bn_train_impl(self, plan, cs, src, s_stride, s_off, h, w, act, a_stride, a_off, pooled, x_input=None):
        B = plan["B"]
        st = _lib.current_stream_ptr()
        z = plan["z:" + cs.name]
        stats = self.bn_stats.data_ptr() + 4 * self.bn_st_off[cs.name]
        slots = self.bn_slots[cs.name]  # one statistic slot per CTA: stored, never accumulated -> no zeroing, bit-reproducible
        bn = cs.bn
        track = bn.track_running_stats and bn.running_mean is not None
        if x_input is None and self.fused_bn_finalize:
            # conv + statistics + BatchNorm finalize in ONE launch: the last CTA of every column block finalizes it
            desc = plan.get("bn:" + cs.name)
            if desc is None:
                li = self.convs.index(cs)
                desc = plan["bn:" + cs.name] = _lib.BnTrain(
                    ptr(bn.weight), ptr(bn.bias), ptr(cs.conv.bias), self._bn(cs, "scale"), self._bn(cs, "shift"),
                    self._bn(cs, "mean"), self._bn(cs, "invstd"), ptr(bn.running_mean) if track else None,
                    ptr(bn.running_var) if track else None, ptr(bn.num_batches_tracked) if track else None,
                    self.bn_fwd_counters.data_ptr() + 4 * 16 * li, float(B * h * w), BN_EPS, BN_MOMENTUM)
            call("b200sr_conv3x3_fwd_bn", ptr(src), s_stride, s_off, cs.cin, self._wp(self.wp_fwd, cs.name), cs.cout,
                 B, h, w, ptr(z), cs.cout, 0, stats, slots, ctypes.byref(desc), st)
        else:
            if x_input is not None:
                call("b200sr_conv1_fwd", ptr(x_input), ptr(cs.conv.weight), None, None, 0, ptr(z), stats, slots, B, h, w, st)
            else:
                call("b200sr_conv3x3_fwd", ptr(src), s_stride, s_off, cs.cin, self._wp(self.wp_fwd, cs.name), cs.cout,
                     B, h, w, ptr(z), cs.cout, 0, None, None, 0, stats, slots, st)
            call("b200sr_bn_finalize", stats, slots, cs.cout, float(B * h * w), ptr(bn.weight), ptr(bn.bias),
                 ptr(cs.conv.bias), BN_EPS, BN_MOMENTUM, self._bn(cs, "scale"), self._bn(cs, "shift"),
                 self._bn(cs, "mean"), self._bn(cs, "invstd"), ptr(bn.running_mean) if track else None,
                 ptr(bn.running_var) if track else None, ptr(bn.num_batches_tracked) if track else None, st)
        call("b200sr_bnrelu_apply", ptr(z), cs.cout, self._bn(cs, "scale"), self._bn(cs, "shift"), ptr(act), a_stride,
             a_off, ptr(pooled), B, h, w, st)

    def forward_train(self, x):
        x = self._check_input(x)
        self.ensure_ready(x.device)
        B, _, H, W = x.shape
        plan = self._plan(B, H, W, True)
        st = _lib.current_stream_ptr()
        fwd_packed = None
        if self.timing_skip_pack and getattr(self, "_packed_once", False):
            self._pack_done = None  # timing experiment only (stale bf16 weights): how much of the packing is exposed?
        elif self.overlap_wgrad:
            fwd_packed = self._repack_staged()
            self._packed_once = True
        else:
            self.repack_weights()
            self._pack_done = None
        ch = self.chans
        cur = None
        for lvl, name in enumerate(["enc1", "enc2", "enc3", "enc4"]):
            c1, c2 = self.blocks[name]
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            a1, cat, pool = plan[f"enc_a1_{lvl}"], plan[f"cat{lvl}"], plan[f"pool{lvl}"]
            if lvl == 2 and fwd_packed is not None:
                torch.cuda.current_stream().wait_event(fwd_packed)  # forward packings of enc3.. are ready
            if lvl == 0:
                self._conv_bn_train(plan, c1, None, 0, 0, h, w, a1, c, 0, None, x_input=x)
            else:
                self._conv_bn_train(plan, c1, cur, c1.cin, 0, h, w, a1, c, 0, None)
            self._conv_bn_train(plan, c2, a1, c, 0, h, w, cat, 2 * c, c, pool)
            cur = pool
        c1, c2 = self.blocks["bottleneck"]
        h, w, c = H >> 4, W >> 4, ch[4]
        self._conv_bn_train(plan, c1, cur, c1.cin, 0, h, w, plan["bot_a1"], c, 0, None)
        self._conv_bn_train(plan, c2, plan["bot_a1"], c, 0, h, w, plan["bot_a2"], c, 0, None)
        cur = plan["bot_a2"]
        for k in (4, 3, 2, 1):
            us = self.ups[k]
            lvl = us.level
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            cat = plan[f"cat{lvl}"]
            call("b200sr_convT2x2_fwd", ptr(cur), us.cin, 0, us.cin, self._wp(self.wp_fwd, us.name), us.cout,
                 ptr(us.mod.bias), B, h // 2, w // 2, ptr(cat), 2 * c, 0, st)
            c1, c2 = self.blocks[f"dec{k}"]
            self._conv_bn_train(plan, c1, cat, 2 * c, 0, h, w, plan[f"dec_a1_{lvl}"], c, 0, None)
            self._conv_bn_train(plan, c2, plan[f"dec_a1_{lvl}"], c, 0, h, w, plan[f"dec_a2_{lvl}"], c, 0, None)
            cur = plan[f"dec_a2_{lvl}"]
        fc = self.head
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        call("b200sr_head_fwd", ptr(cur), ptr(fc.weight), ptr(fc.bias), ptr(out), B * H * W, st)
        self._saved = (plan, x)
        return out

    # ------------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------------
    def _bn_bwd(self, plan, cs, dy, dy_stride, dy_off, h, w, dz, sums_ready=False):
        """BatchNorm+ReLU backward of one layer: dy -> dz (dense), dgamma/dbeta into the flat gradient. sums_ready: pass 1
        (the reduction) was already done by the kernel that produced dy (fused max-pool backward)."""
        B = plan["B"]
        st = _lib.current_stream_ptr()
        z = plan["z:" + cs.name]
        sums = self.bn_sums.data_ptr() + 4 * self.bn_sum_off[cs.name]
        npix = B * h * w
        sc, sh, mu, iv = (self._bn(cs, k) for k in ("scale", "shift", "mean", "invstd"))
        if not sums_ready:
            call("b200sr_bn_bwd_reduce_det", dy, dy_stride, dy_off, ptr(z), cs.cout, sc, sh, mu, iv, sums, ptr(self.red_ws),
                 self.red_ws.numel(), ptr(self.red_counters), None, npix, st)
        g = self.flat_g.data_ptr()
        call("b200sr_bn_bwd_apply_fused", dy, dy_stride, dy_off, ptr(z), cs.cout, sc, sh, mu, iv, sums, 1,
             float(npix), g + 4 * self.off_of[id(cs.bn.weight)], g + 4 * self.off_of[id(cs.bn.bias)], dz, npix, st)

    def _g(self, param):
        """Device address of a parameter's gradient inside the flat gradient buffer."""
        return self.flat_g.data_ptr() + 4 * self.off_of[id(param)]

    def backward(self, dout, bucket_hook=None, want_dx=False):
        """See _backward. With overlap enabled the data-gradient chain runs on a HIGH-priority stream and the weight
        gradients on a default (lowest) priority stream: the chain is the critical path, the wgrad kernels only fill
        the SMs it leaves idle (during the bandwidth-bound BatchNorm / pooling kernels)."""
        if not (self.overlap_wgrad and self.hp_chain):
            return self._backward(dout, bucket_hook, want_dx)
        cur = torch.cuda.current_stream()
        if self._hp is None:
            self._hp = torch.cuda.Stream(device=self.device, priority=-1)
        self._hp.wait_stream(cur)
        with torch.cuda.stream(self._hp):
            out = self._backward(dout, bucket_hook, want_dx)
        cur.wait_stream(self._hp)
        return out

    def _backward(self, dout, bucket_hook=None, want_dx=False):
        """Full backward of the last train-mode forward. dout: (B,1,H,W) fp32.
        Gradients land in self.flat_g (views: self.grad_views, in model.parameters() order).
        bucket_hook(lo, hi), if given, is called as soon as flat_g[lo:hi] is final (reverse forward order, one range
        per block: final_conv+dec1+upconv1, dec2+upconv2, ..., bottleneck, enc4, ..., enc1).

        Two streams: the data-gradient chain (BN backward -> dgrad -> ...) runs on the current stream; every weight
        gradient (tensor-core split-K kernels, their fixed-order reduction into the parameter layout and the bucket hook)
        runs on a side stream, ordered by events, so the tensor-bound wgrad kernels overlap the bandwidth-bound BatchNorm /
        pooling backward kernels of the following layers. Each layer has its own dz buffer, hence no write-after-read
        hazard between the streams inside a step; the streams are joined at the end.

        Every reduction is deterministic (per-CTA / per-split partials, fixed-order second stage; include/b200sr.h):
        the gradients are WRITTEN, nothing is zeroed per step, and two runs on the same inputs give the same bits."""
        if self._saved is None:
            raise _lib.B200SRError("backward() without a preceding train-mode forward")
        plan, x = self._saved
        B, H, W = plan["B"], plan["H"], plan["W"]
        self.dx_input = None  # gradient w.r.t. the network input (B,2,H,W) fp32, filled when want_dx
        main = torch.cuda.current_stream()
        if getattr(self, "_pack_done", None) is not None:
            main.wait_event(self._pack_done)  # dgrad packings were produced on the side stream during forward
            self._pack_done = None
        st = main.cuda_stream
        overlap = self.overlap_wgrad
        if overlap:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
        else:
            side = main
        sst = side.cuda_stream
        ch = self.chans
        dout = dout.contiguous().float()
        s0, s1, s2 = (t.data_ptr() for t in plan["scratch"])
        fc = self.head
        ev_i = [0]
        wg_ws, wg_n = ptr(self.wg_ws), self.wg_ws.numel()
        n_slots = self.n_sms

        def side_after_main():
            """Everything enqueued on the main stream so far happens-before what is enqueued on the side stream next."""
            if not overlap:
                return
            if ev_i[0] == len(self._events):
                self._events.append(torch.cuda.Event())
            ev = self._events[ev_i[0]]
            ev_i[0] += 1
            ev.record(main)
            side.wait_event(ev)

        def report(lo_param, hi_off):
            """flat_g[off(lo_param) : hi_off] is final once everything enqueued so far (both streams) has run."""
            lo = self.off_of[id(lo_param)]
            if bucket_hook is not None:
                side_after_main()  # BN / bias gradients of the range are produced on the main stream
                with torch.cuda.stream(side):
                    bucket_hook(lo, hi_off)
            return lo

        def wgrad3(src, s_stride, cin, dz, cout, h, w, weight):
            call("b200sr_conv3x3_wgrad_det", src, s_stride, 0, cin, dz, cout, 0, cout, B, h, w, self._g(weight), cin, 0,
                 wg_ws, wg_n, sst)

        # head
        a_last = plan["dec_a2_0"]
        head_sums_ready = self.fused_head_bnred
        if head_sums_ready:
            # 1x1 head backward + pass 1 of the BatchNorm backward of dec1.conv.3 (whose gradient it writes) in one kernel
            cs = self.blocks["dec1"][1]
            call("b200sr_head_bwd_bnred", ptr(dout), ptr(a_last), ptr(fc.weight), s0, self._g(fc.weight), self._g(fc.bias),
                 ptr(plan["z:" + cs.name]), self._bn(cs, "scale"), self._bn(cs, "shift"), self._bn(cs, "mean"),
                 self._bn(cs, "invstd"), self.bn_sums.data_ptr() + 4 * self.bn_sum_off[cs.name], B * H * W,
                 ptr(self.red_ws), self.red_ws.numel(), ptr(self.red_counters) + 4 * 60, st)
        else:
            call("b200sr_head_bwd_det", ptr(dout), ptr(a_last), ptr(fc.weight), s0, self._g(fc.weight), self._g(fc.bias),
                 B * H * W, ptr(self.red_ws), self.red_ws.numel(), ptr(self.red_counters) + 4 * 63, st)
        dy = s0  # gradient w.r.t. the current block's output activation (dense)

        def block_bwd(name, lvl, in_buf, in_stride, in_c, dx_dst, dx_stride, dx_stats, dy_ptr, x_input=None,
                      sums_ready=False):
            with _Nvtx("bwd/" + name):
                return block_bwd_impl(name, lvl, in_buf, in_stride, in_c, dx_dst, dx_stride, dx_stats, dy_ptr, x_input,
                                      sums_ready)

        def block_bwd_impl(name, lvl, in_buf, in_stride, in_c, dx_dst, dx_stride, dx_stats, dy_ptr, x_input=None,
                           sums_ready=False):
            """Backward through a UNetBlock: dy (dense, cout ch) -> dx into dx_dst (in_c channels)."""
            c1, c2 = self.blocks[name]
            h, w, c = H >> lvl, W >> lvl, c2.cout
            dy1 = [p for p in (s0, s1, s2) if p != dy_ptr and p != dx_dst][0]
            dz2, dz1 = ptr(plan["dz:" + c2.name]), ptr(plan["dz:" + c1.name])
            a1 = plan["bot_a1"] if name == "bottleneck" else plan[f"{name[:3]}_a1_{lvl}"]
            # Stream choreography: the persistent dgrad and wgrad kernels each want every SM (one CTA per SM, > 200 KB of
            # shared memory); launched together they serialise. `late` (experiment, off) orders the wgrad behind the dgrad.
            late = self.wgrad_late
            self._bn_bwd(plan, c2, dy_ptr, c, 0, h, w, dz2, sums_ready=sums_ready)
            if not late:
                side_after_main()
                wgrad3(ptr(a1), c, c, dz2, c, h, w, c2.conv.weight)
            call("b200sr_conv3x3_dgrad", dz2, c, 0, c, self._wp(self.wp_dgrad, c2.name), c, B, h, w, dy1, c, 0, None,
                 0, st)
            if late:
                side_after_main()
                wgrad3(ptr(a1), c, c, dz2, c, h, w, c2.conv.weight)
            self._bn_bwd(plan, c1, dy1, c, 0, h, w, dz1)
            if x_input is not None:
                side_after_main()
                call("b200sr_conv1_wgrad_det", ptr(x_input), dz1, self._g(c1.conv.weight), B, h, w, wg_ws, wg_n, sst)
                if want_dx:
                    self.dx_input = torch.empty((B, 2, h, w), dtype=torch.float32, device=x_input.device)
                    call("b200sr_conv1_dgrad", dz1, ptr(c1.conv.weight), ptr(self.dx_input), B, h, w, st)
                return None
            if not late:
                side_after_main()
                wgrad3(ptr(in_buf), in_stride, in_c, dz1, c, h, w, c1.conv.weight)
            if dx_stats and not self._ab_full_stats:
                # decoder conv.0: the gradient of the concat buffer; the column sums of its upsampled half (first in_c/2
                # channels) are the ConvTranspose2d bias gradient
                call("b200sr_conv3x3_dgrad_colsum", dz1, c, 0, c, self._wp(self.wp_dgrad, c1.name), in_c, B, h, w, dx_dst,
                     dx_stride, 0, dx_stats, n_slots, in_c // 2, st)
            else:
                call("b200sr_conv3x3_dgrad", dz1, c, 0, c, self._wp(self.wp_dgrad, c1.name), in_c, B, h, w, dx_dst,
                     dx_stride, 0, dx_stats, n_slots if dx_stats else 0, st)
            if late:
                side_after_main()
                wgrad3(ptr(in_buf), in_stride, in_c, dz1, c, h, w, c1.conv.weight)
            return dx_dst

        hi = self.p_total
        # decoder, shallow -> deep
        for k in (1, 2, 3, 4):
            us = self.ups[k]
            lvl = us.level
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            dcat = plan[f"dcat{lvl}"]
            up_stats = self.up_stats.data_ptr() + 4 * self.up_st_off[k]
            block_bwd(f"dec{k}", lvl, plan[f"cat{lvl}"], 2 * c, 2 * c, ptr(dcat), 2 * c, up_stats, dy,
                      sums_ready=(k == 1 and head_sums_ready))
            # ConvTranspose bias gradient = column sums of the upsampled half of dcat: per-CTA slots [sum | sumsq][2c]
            # from the dgrad epilogue, added in slot order
            call("b200sr_sum_slots", up_stats, n_slots, 2 * 2 * c, c, self._g(us.mod.bias), st)
            x_up = plan["bot_a2"] if k == 4 else plan[f"dec_a2_{lvl + 1}"]

            def up_wgrad():
                side_after_main()
                call("b200sr_convT2x2_wgrad_det", ptr(dcat), 2 * c, 0, c, ptr(x_up), us.cin, 0, us.cin, B, h // 2, w // 2,
                     self._g(us.mod.weight), wg_ws, wg_n, sst)

            if not self.wgrad_late:
                up_wgrad()
            dy = s0 if dy != s0 else s1
            call("b200sr_convT2x2_dgrad", ptr(dcat), 2 * c, 0, c, self._wp(self.wp_dgrad, us.name), us.cin, B, h // 2,
                 w // 2, dy, us.cin, 0, st)
            if self.wgrad_late:
                up_wgrad()
            hi = report(us.mod.weight, hi)  # upconv_k is the lowest flat offset of {upconv_k, dec_k[, final_conv]}

        # bottleneck: dx = gradient w.r.t. pool3 (dense)
        dpool = [p for p in (s0, s1, s2) if p != dy][0]
        block_bwd("bottleneck", 4, plan["pool3"], ch[3], ch[3], dpool, ch[3], None, dy)
        hi = report(self.blocks["bottleneck"][0].conv.weight, hi)

        # encoder, deep -> shallow
        for lvl in (3, 2, 1, 0):
            name = f"enc{lvl + 1}"
            h, w, c = H >> lvl, W >> lvl, ch[lvl]
            cat, dcat = plan[f"cat{lvl}"], plan[f"dcat{lvl}"]
            dy = [p for p in (s0, s1, s2) if p != dpool][0]
            fused = self.fused_pool_bnred
            if fused:
                # max-pool backward + skip add + pass 1 of the BatchNorm backward of this block's conv.3 in one launch
                c2 = self.blocks[name][1]
                call("b200sr_maxpool2x2_bwd_bnred", ptr(cat), 2 * c, c, dpool, ptr(dcat), 2 * c, c, c, dy,
                     ptr(plan["z:" + c2.name]), self._bn(c2, "scale"), self._bn(c2, "shift"), self._bn(c2, "mean"),
                     self._bn(c2, "invstd"), self.bn_sums.data_ptr() + 4 * self.bn_sum_off[c2.name], ptr(self.red_ws),
                     self.red_ws.numel(), ptr(self.red_counters), B, h, w, st)
            else:
                call("b200sr_maxpool2x2_bwd", ptr(cat), 2 * c, c, dpool, ptr(dcat), 2 * c, c, c, dy, B, h, w, st)
            if lvl == 0:
                block_bwd(name, lvl, None, 0, 2, None, 0, None, dy, x_input=x, sums_ready=fused)
            else:
                nxt = [p for p in (s0, s1, s2) if p != dy][0]
                block_bwd(name, lvl, plan[f"pool{lvl - 1}"], ch[lvl - 1], ch[lvl - 1], nxt, ch[lvl - 1], None, dy,
                          sums_ready=fused)
                dpool = nxt
            hi = report(self.blocks[name][0].conv.weight, hi)
        if overlap:
            main.wait_stream(side)
        self._saved = None
        return self.grad_views

    def _params_index(self, param):
        if not hasattr(self, "_pindex"):
            self._pindex = {id(p): i for i, p in enumerate(self._params())}
        return self._pindex[id(param)]
