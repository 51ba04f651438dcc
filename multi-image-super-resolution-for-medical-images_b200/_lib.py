"""ctypes binding of libb200sr.so (C ABI declared in include/b200sr.h).

The library is the product: there is no Python/torch fallback for any op. If the shared object is missing the
import of this module still succeeds (so CPU-only host logic such as state_dict handling keeps working), but the
first call of any op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SR_LIB") or os.path.join(_HERE, "libb200sr.so")  # B200SR_LIB: A/B builds of the library

_lib = None
_load_error = None

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_P = c_void_p
_SIGNATURES = {
    "b200sr_version": [],
    "b200sr_last_error": [],
    "b200sr_last_wgrad_launches": [],
    "b200sr_device_ok": [],
    "b200sr_conv3x3_fwd": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, c_int,
                           _P, c_int, _P],
    "b200sr_conv3x3_fwd_bn": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, _P],
    "b200sr_conv3x3_dgrad": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int,
                             _P],
    "b200sr_conv3x3_dgrad_colsum": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int,
                                    c_int, _P],
    "b200sr_conv3x3_dgrad_relu": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int,
                                  c_int, _P, c_int, _P],
    "b200sr_convT2x2_fwd": [_P, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "b200sr_convT2x2_dgrad": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "b200sr_conv3x3_wgrad": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "b200sr_convT2x2_wgrad": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "b200sr_pack_jobs": [_P, c_int, _P],
    "b200sr_bn_fold_eval": [_P, c_int, c_float, _P],
    "b200sr_conv1_fwd": [_P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_conv1_dgrad": [_P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_conv1_wgrad": [_P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_bn_finalize": [_P, c_int, c_int, c_double, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P, _P],
    "b200sr_bnrelu_apply": [_P, c_int, _P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_bn_train_apply": [_P, c_int, _P, c_int, c_double, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P,
                              c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_bn_bwd_apply_fused": [_P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, c_int, c_double, _P, _P, _P, c_int64,
                                  _P],
    "b200sr_maxpool2x2_fwd": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_maxpool2x2_bwd": [_P, c_int, c_int, _P, _P, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_bn_bwd_reduce": [_P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, c_int, c_int64, _P],
    "b200sr_bn_bwd_finalize": [_P, c_int, c_int, c_double, _P, _P, _P, _P, _P],
    "b200sr_bn_bwd_apply": [_P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_int64, _P],
    "b200sr_relu_bwd": [_P, _P, _P, c_int64, _P],
    "b200sr_feat_mse_grad": [_P, _P, _P, _P, c_float, c_int64, _P],
    "b200sr_conv7_fwd": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_conv7_wgrad": [_P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_maxpool3x3_fwd": [_P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_maxpool3x3_bwd": [_P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_bn_add_relu": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int64, _P],
    "b200sr_bn_bwd_masked": [_P, _P, _P, c_int, _P, _P, _P, _P, _P, c_int, c_double, _P, _P, _P, c_int64, _P],
    "b200sr_add_masked": [_P, _P, _P, _P, c_int64, _P],
    "b200sr_headw_fwd": [_P, c_int, _P, _P, _P, c_int64, _P],
    "b200sr_headw_bwd": [_P, _P, c_int, _P, _P, _P, _P, c_int64, _P],
    "b200sr_conv1x1": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, c_int, _P],
    "b200sr_conv1x1_wgrad": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "b200sr_fd_time_mlp_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, c_int, _P],
    "b200sr_fd_time_bias": [_P, _P, _P, _P, c_int, _P],
    "b200sr_fd_convin_fwd": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_convin_wgrad": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_relu_bwd_bias": [_P, c_int, c_int, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_fd_upsample2x_bwd_relu": [_P, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_fd_head_bwd_relu": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_maxpool2x2_bwd_relu": [_P, c_int, c_int, _P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_bias_finish": [_P, c_int, c_int, _P],
    "b200sr_fd_time_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_upsample2x_fwd": [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "b200sr_fd_upsample2x_bwd": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_q_sample": [_P, _P, _P, _P, c_int, c_int, c_int, _P],
    "b200sr_fd_ddim_update": [_P, _P, c_float, c_float, c_int, c_int64, _P],
    "b200sr_grad_clip": [_P, c_int64, _P, c_float, c_float, _P],
    "b200sr_head_fwd": [_P, _P, _P, _P, c_int64, _P],
    "b200sr_head_bwd": [_P, _P, _P, _P, _P, _P, c_int64, _P],
    "b200sr_mse_ssim": [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_int, c_float, c_float, c_float, c_float, c_float,
                        _P],
    "b200sr_adam_step": [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_int64, c_float, _P],
    "b200sr_adam_step_dev": [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, _P, c_float, _P],
    # deterministic (bit-reproducible) reductions
    "b200sr_conv3x3_wgrad_det": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int,
                                 _P, c_int64, _P],
    "b200sr_convT2x2_wgrad_det": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int64,
                                  _P],
    "b200sr_conv1x1_wgrad_det": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int64,
                                 _P],
    "b200sr_conv1_wgrad_det": [_P, _P, _P, c_int, c_int, c_int, _P, c_int64, _P],
    "b200sr_sum_slots": [_P, c_int, c_int64, c_int, _P, _P],
    "b200sr_bn_bwd_ws_floats": [c_int],
    "b200sr_bn_bwd_reduce_det": [_P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, c_int64, _P],
    "b200sr_maxpool2x2_bwd_bnred": [_P, c_int, c_int, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64,
                                    _P, c_int, c_int, c_int, _P],
    "b200sr_head_bwd_det": [_P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P],
    "b200sr_head_bwd_bnred": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P],
    "b200sr_mse_ssim_det": [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_int, c_float, c_float, c_float, c_float, c_float,
                            _P, c_int64, _P, _P],
    "b200sr_adam_step_auto": [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, _P, c_float, _P],
    # fp32-accuracy eval mode (bf16x3 operand split)
    "b200sr_conv3x3_fwd_split": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P, _P,
                                 c_int, _P],
    "b200sr_convT2x2_fwd_split": [_P, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_conv1_fwd_split": [_P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_maxpool2x2_fwd_split": [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P],
    "b200sr_head_fwd_split": [_P, _P, _P, _P, c_int64, _P],
    "b200sr_volume_metrics": [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int64, _P, _P],
    "b200sr_nchw_f32_to_nhwc_bf16": [_P, _P, c_int, c_int, c_int, c_int, _P],
    "b200sr_nhwc_bf16_to_nchw_f32": [_P, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P],
}
_RESTYPES = {"b200sr_last_error": c_char_p, "b200sr_bn_bwd_ws_floats": c_int64}
_PLAIN_VALUE = ("b200sr_version", "b200sr_last_error", "b200sr_bn_bwd_ws_floats", "b200sr_last_wgrad_launches")  # return a value, not a status code

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class B200SRError(RuntimeError):
    pass


class BnTrain(ctypes.Structure):
    """b200sr_bn_train (include/b200sr.h): host struct of device pointers for the fused conv + BatchNorm finalize."""
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p), ("conv_bias", c_void_p), ("scale", c_void_p),
                ("shift", c_void_p), ("save_mean", c_void_p), ("save_invstd", c_void_p), ("running_mean", c_void_p),
                ("running_var", c_void_p), ("num_batches_tracked", c_void_p), ("counters", c_void_p),
                ("count", c_double), ("eps", c_float), ("momentum", c_float)]


def load():
    """Load the shared library (once). Raises B200SRError if it has not been built."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        _load_error = (f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU/torch fallback for the b200sr hot path.")
        raise B200SRError(_load_error)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().b200sr_last_error()
    return msg.decode() if msg else ""


# ---- launch accounting and optional per-call CUDA-event timing (used by bench.py for the roofline) ------------
LAUNCH_COUNTER = {"n": 0}
# entry points that enqueue more than one kernel (split-K gradient + its fixed-order reduction, multi-kernel helpers)
_WGRAD_DET = ("b200sr_conv3x3_wgrad_det", "b200sr_convT2x2_wgrad_det", "b200sr_conv1x1_wgrad_det")  # 2 or 3: asked per call
_KERNELS_PER_CALL = {"b200sr_conv1_wgrad_det": 2, "b200sr_bn_bwd_masked": 2, "b200sr_grad_clip": 2, "b200sr_volume_metrics": 3, "b200sr_fd_time_bwd": 5,
                     "b200sr_bn_bwd_ws_floats": 0, "b200sr_version": 0, "b200sr_device_ok": 0, "b200sr_last_wgrad_launches": 0}
GEMM_OPS = ("b200sr_conv3x3_fwd", "b200sr_conv3x3_dgrad", "b200sr_conv3x3_dgrad_relu", "b200sr_conv3x3_wgrad", "b200sr_convT2x2_fwd",
            "b200sr_convT2x2_dgrad", "b200sr_convT2x2_wgrad", "b200sr_conv1x1", "b200sr_conv1x1_wgrad",
            "b200sr_conv3x3_wgrad_det", "b200sr_convT2x2_wgrad_det", "b200sr_conv1x1_wgrad_det", "b200sr_conv3x3_fwd_bn", "b200sr_conv3x3_dgrad_colsum", "b200sr_conv3x3_fwd_split", "b200sr_convT2x2_fwd_split")
_profile = None  # list of (name, start_event, end_event, flop, bytes) while profiling is enabled


def _cost(name, a):
    """Algorithmic (flop, bytes) of one call, from its argument list (see include/b200sr.h for the order)."""
    if name in ("b200sr_conv3x3_fwd", "b200sr_conv3x3_dgrad", "b200sr_conv3x3_dgrad_relu", "b200sr_conv3x3_fwd_split",
                "b200sr_conv3x3_fwd_bn", "b200sr_conv3x3_dgrad_colsum"):
        return 2.0 * a[6] * a[7] * a[8] * a[3] * a[5] * 9, 0.0   # (split: issued flops, 3x the algorithmic ones)
    if name in ("b200sr_convT2x2_dgrad",):
        return 2.0 * a[6] * a[7] * a[8] * a[3] * a[5] * 4, 0.0
    if name == "b200sr_convT2x2_fwd":
        return 2.0 * a[7] * a[8] * a[9] * a[3] * a[5] * 4, 0.0
    if name in ("b200sr_conv3x3_wgrad", "b200sr_conv3x3_wgrad_det"):
        return 2.0 * a[8] * a[9] * a[10] * a[3] * a[7] * 9, 0.0
    if name in ("b200sr_convT2x2_wgrad", "b200sr_convT2x2_wgrad_det"):
        return 2.0 * a[8] * a[9] * a[10] * a[3] * a[7] * 4, 0.0
    if name == "b200sr_conv1x1_wgrad_det":
        return 2.0 * a[8] * a[9] * a[10] * a[3] * a[7], 0.0
    if name == "b200sr_maxpool2x2_bwd_bnred":
        return 0.0, a[18] * a[19] * a[20] * a[7] * 2.0 * 4.25   # act, dskip, z read + dy written + dpool/4
    if name == "b200sr_bn_bwd_reduce_det":
        return 0.0, a[14] * a[4] * 2.0 * 2
    if name == "b200sr_head_bwd_det":
        return 0.0, a[6] * (4.0 + 128 + 128)
    if name == "b200sr_head_bwd_bnred":
        return 0.0, a[12] * (4.0 + 128 + 128 + 128)   # dout, act and z read, dact written
    if name == "b200sr_mse_ssim_det":
        return 0.0, a[4] * a[5] * a[6] * 4.0 * (3 if a[2] else 2)
    if name == "b200sr_adam_step_auto":
        return 0.0, a[4] * 4.0 * 7
    if name == "b200sr_conv1_wgrad_det":
        return 0.0, a[3] * a[4] * a[5] * (8.0 + 128)
    if name == "b200sr_conv1x1":
        return 2.0 * a[6] * a[7] * a[8] * a[3] * a[5], 0.0
    if name == "b200sr_conv1x1_wgrad":
        return 2.0 * a[8] * a[9] * a[10] * a[3] * a[7], 0.0
    if name == "b200sr_bnrelu_apply":
        n = a[8] * a[9] * a[10] * a[1] * 2.0
        return 0.0, n * (2.25 if a[7] else 2.0)
    if name == "b200sr_bn_train_apply":
        n = a[20] * a[21] * a[22] * a[1] * 2.0
        return 0.0, n * (2.25 if a[19] else 2.0)
    if name == "b200sr_bn_bwd_apply_fused":
        return 0.0, a[15] * a[4] * 2.0 * 3
    if name == "b200sr_maxpool2x2_fwd":
        return 0.0, a[5] * a[6] * a[7] * a[3] * 2.0 * 1.25
    if name == "b200sr_maxpool2x2_bwd":
        return 0.0, a[9] * a[10] * a[11] * a[7] * 2.0 * (3.25 if a[4] else 2.25)
    if name == "b200sr_bn_bwd_reduce":
        return 0.0, a[11] * a[4] * 2.0 * 2
    if name == "b200sr_bn_bwd_apply":
        return 0.0, a[12] * a[4] * 2.0 * 3
    if name == "b200sr_relu_bwd":
        return 0.0, a[3] * 2.0 * 3
    if name == "b200sr_feat_mse_grad":
        return 0.0, a[5] * 2.0 * (3 if a[2] else 2)
    if name == "b200sr_head_fwd":
        return 0.0, a[4] * (128.0 + 4)
    if name == "b200sr_head_bwd":
        return 0.0, a[6] * (4.0 + 128 + 128)
    if name == "b200sr_mse_ssim":
        return 0.0, a[4] * a[5] * a[6] * 4.0 * (3 if a[2] else 2)
    if name in ("b200sr_adam_step", "b200sr_adam_step_dev"):
        return 0.0, a[4] * 4.0 * 7
    if name == "b200sr_conv1_fwd":
        return 0.0, a[8] * a[9] * a[10] * (8.0 + 128)
    if name == "b200sr_conv1_dgrad":
        return 0.0, a[3] * a[4] * a[5] * (8.0 + 128)
    if name == "b200sr_conv1_wgrad":
        return 0.0, a[3] * a[4] * a[5] * (8.0 + 128)
    return 0.0, 0.0


def enable_profiling(on: bool):
    global _profile
    _profile = [] if on else None
    return _profile


def collect_profile():
    """Aggregate the recorded calls: {op: {"ms", "n", "flop", "bytes"}} (synchronises)."""
    import torch
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1, flop, nbytes in _profile or []:
        d = agg.setdefault(name, {"ms": 0.0, "n": 0, "flop": 0.0, "bytes": 0.0})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
        d["flop"] += flop
        d["bytes"] += nbytes
    return agg


def call(name: str, *args):
    """Call an exported op; non-zero return codes become B200SRError(last_error)."""
    fn = getattr(load(), name)
    if _profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _profile.append((name, e0, e1) + _cost(name, args))
    else:
        rc = fn(*args)
    if name in _WGRAD_DET:
        LAUNCH_COUNTER["n"] += load().b200sr_last_wgrad_launches() if rc == 0 else 0
    else:
        LAUNCH_COUNTER["n"] += _KERNELS_PER_CALL.get(name, 1)
    if name in _PLAIN_VALUE:
        return rc
    if rc != 0:
        raise B200SRError(f"{name} failed (code {rc}): {last_error()}")


def ptr(t, elem_offset: int = 0):
    """Device pointer of a torch tensor (or None -> NULL), optionally advanced by elements."""
    if t is None:
        return None
    return t.data_ptr() + elem_offset * t.element_size()


def current_stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_device():
    """Raise unless the current CUDA device is a compute-capability 10.x part."""
    import torch
    if not torch.cuda.is_available():
        raise B200SRError("b200sr needs a CUDA sm_100a device; no CUDA device is available and there is no CPU path")
    call("b200sr_device_ok")
