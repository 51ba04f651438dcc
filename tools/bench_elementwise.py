"""Bandwidth kernels at the UNet's real shapes (B=32): algorithmic GB/s per kernel and level."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200sr
from b200sr import _lib
from b200sr._lib import call, ptr

B = int(os.environ.get("B", "32"))
dev = "cuda"
st = _lib.current_stream_ptr()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


tot = {}
for hw, C in ((256, 64), (128, 128), (64, 256), (32, 512), (16, 1024)):
    n = B * hw * hw
    z = torch.randn(B, hw, hw, C, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, hw, hw, C, device=dev).to(torch.bfloat16)
    dz = torch.empty_like(z)
    act = torch.empty(B, hw, hw, 2 * C, dtype=torch.bfloat16, device=dev)
    pooled = torch.empty(B, hw // 2, hw // 2, C, dtype=torch.bfloat16, device=dev)
    dpool = torch.randn(B, hw // 2, hw // 2, C, device=dev).to(torch.bfloat16)
    p = [torch.rand(C, device=dev) + 0.5 for _ in range(6)]
    sums = torch.zeros(16, 2, C, device=dev)
    by = n * C * 2.0
    ks = {
        "bn_bwd_reduce": (lambda: call("b200sr_bn_bwd_reduce", ptr(dy), C, 0, ptr(z), C, ptr(p[0]), ptr(p[1]), ptr(p[2]), ptr(p[3]), ptr(sums), 16, n, st), 2 * by),
        "bn_bwd_apply": (lambda: call("b200sr_bn_bwd_apply", ptr(dy), C, 0, ptr(z), C, ptr(p[0]), ptr(p[1]), ptr(p[2]), ptr(p[3]), ptr(p[4]), ptr(p[5]), ptr(dz), n, st), 3 * by),
        "bnrelu_apply_pool": (lambda: call("b200sr_bnrelu_apply", ptr(z), C, ptr(p[0]), ptr(p[1]), ptr(act), 2 * C, C, ptr(pooled), B, hw, hw, st), 2.25 * by),
        "bnrelu_apply": (lambda: call("b200sr_bnrelu_apply", ptr(z), C, ptr(p[0]), ptr(p[1]), ptr(dz), C, 0, None, B, hw, hw, st), 2 * by),
        "maxpool_bwd": (lambda: call("b200sr_maxpool2x2_bwd", ptr(act), 2 * C, C, ptr(dpool), ptr(act), 2 * C, 0, C, ptr(dz), B, hw, hw, st), 3.25 * by),
    }
    line = f"{hw:3d}^2 C={C:4d} |"
    for k, (fn, nbytes) in ks.items():
        ms = timeit(fn)
        t = tot.setdefault(k, [0.0, 0.0]); t[0] += ms; t[1] += nbytes
        line += f" {k} {ms*1e3:6.1f}us {nbytes/ms/1e6:5.0f}GB/s |"
    print(line, flush=True)
for k, (ms, nb) in tot.items():
    print(f"total {k}: {ms:.3f} ms {nb/ms/1e6:.0f} GB/s")
# loss kernel at two batch sizes
crit = b200sr.CombinedLoss(1.0, 0.005)
for bb in (32, 512):
    x = torch.randn(bb, 1, 256, 256, device=dev); y = torch.randn(bb, 1, 256, 256, device=dev)
    ms = timeit(lambda: crit.value_and_grad(x, y))
    print(f"mse_ssim B={bb}: {ms*1e3:.1f} us  {bb*65536*12/ms/1e6:.0f} GB/s algorithmic")
n = 31_043_000
pp = [torch.zeros(n, device=dev) for _ in range(4)]
ms = timeit(lambda: call("b200sr_adam_step", ptr(pp[0]), ptr(pp[1]), ptr(pp[2]), ptr(pp[3]), n, 1e-4, 0.9, 0.999, 1e-8, 1, 1.0, st))
print(f"adam: {ms*1e3:.1f} us {n*28/ms/1e6:.0f} GB/s")
