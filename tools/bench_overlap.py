"""Does a bandwidth-bound BatchNorm-backward pass overlap a tensor-core wgrad / dgrad kernel running on another
stream? Prints alone-times and the concurrent makespan for a few layer shapes (B200, B=32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200sr
from b200sr import _lib
from b200sr._lib import call, ptr

B = int(os.environ.get("B", "32"))
dev = "cuda"
s_main, s_side = torch.cuda.Stream(), torch.cuda.Stream()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def span(fns, reps=7):
    """fns: list of (stream, callable). Median makespan in ms, L2 flushed before each rep."""
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        for s, _ in fns:
            s.wait_stream(cur)
        for s, f in fns:
            with torch.cuda.stream(s):
                f(s.cuda_stream)
        for s, _ in fns:
            cur.wait_stream(s)
        e1.record(cur)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


for name, hw, cin, cout in (("dec1.0", 256, 128, 64), ("enc2.3", 128, 128, 128), ("dec3.0", 64, 512, 256)):
    x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    wf = (torch.randn(cout * 9 * cin, device=dev) * 0.02).to(torch.bfloat16)
    dx = torch.empty(B, hw, hw, cin, dtype=torch.bfloat16, device=dev)
    G = torch.zeros(9 * cin * cout, device=dev)
    # an independent BatchNorm backward (reduce + apply) of the same spatial size, cout channels
    dy2 = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    z2 = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    dz2 = torch.empty_like(z2)
    vec = [torch.rand(cout, device=dev) + 0.5 for _ in range(4)]
    sums = torch.zeros(4, 2, cout, device=dev)
    dgb = torch.zeros(2, cout, device=dev)
    npix = B * hw * hw

    def wgrad(st):
        call("b200sr_conv3x3_wgrad", ptr(x), cin, 0, cin, ptr(dz), cout, 0, cout, B, hw, hw, ptr(G), st)

    def dgrad(st):
        call("b200sr_conv3x3_dgrad", ptr(dz), cout, 0, cout, ptr(wf), cin, B, hw, hw, ptr(dx), cin, 0, None, 0, st)

    def bn(st):
        call("b200sr_bn_bwd_reduce", ptr(dy2), cout, 0, ptr(z2), cout, *[ptr(v) for v in vec], ptr(sums), 4, npix, st)
        call("b200sr_bn_bwd_apply_fused", ptr(dy2), cout, 0, ptr(z2), cout, *[ptr(v) for v in vec], ptr(sums), 4,
             float(npix), ptr(dgb[0]), ptr(dgb[1]), ptr(dz2), npix, st)

    t_w, t_d, t_b = span([(s_main, wgrad)]), span([(s_main, dgrad)]), span([(s_main, bn)])
    t_wb = span([(s_side, wgrad), (s_main, bn)])
    t_bw = span([(s_main, bn), (s_side, wgrad)])
    t_db = span([(s_side, dgrad), (s_main, bn)])
    t_wd = span([(s_side, wgrad), (s_main, dgrad)])
    print(f"{name}: wgrad {t_w*1e3:.0f} us, dgrad {t_d*1e3:.0f} us, bn_bwd {t_b*1e3:.0f} us | wgrad||bn {t_wb*1e3:.0f} "
          f"(bn first: {t_bw*1e3:.0f}; sum {1e3*(t_w+t_b):.0f}) | dgrad||bn {t_db*1e3:.0f} (sum {1e3*(t_d+t_b):.0f}) | "
          f"wgrad||dgrad {t_wd*1e3:.0f} (sum {1e3*(t_w+t_d):.0f})", flush=True)
