"""Per-layer timing of the tensor-core kernels at the UNet's real shapes (B per GPU = 32 by default)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200sr
from b200sr import _lib
from b200sr._lib import call, ptr

B = int(os.environ.get("B", "32"))
REPS = 10
dev = "cuda"
st = _lib.current_stream_ptr()
layers = [("enc1.3", 256, 64, 64), ("enc2.0", 128, 64, 128), ("enc2.3", 128, 128, 128), ("enc3.0", 64, 128, 256),
          ("enc3.3", 64, 256, 256), ("enc4.0", 32, 256, 512), ("enc4.3", 32, 512, 512), ("bott.0", 16, 512, 1024),
          ("bott.3", 16, 1024, 1024), ("dec4.0", 32, 1024, 512), ("dec3.0", 64, 512, 256), ("dec2.0", 128, 256, 128),
          ("dec1.0", 256, 128, 64)]
which = [a for a in sys.argv[1:] if a != "convT"] or (["fwd", "dgrad", "wgrad"] if len(sys.argv) == 1 else [])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


tot = {k: [0.0, 0.0] for k in which}
print(f"B={B}  (median of {REPS}, L2 flushed between reps)")
for name, hw, cin, cout in (layers if which else []):
    x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    wf = (torch.randn(cout * 9 * cin, device=dev) * 0.02).to(torch.bfloat16)
    out = torch.empty(B, hw, hw, cout, dtype=torch.bfloat16, device=dev)
    dx = torch.empty(B, hw, hw, cin, dtype=torch.bfloat16, device=dev)
    G = torch.zeros(9 * cin * cout, device=dev)
    stats = torch.zeros(16, 2, cout, device=dev)
    flop = 2.0 * B * hw * hw * cin * cout * 9
    fns = {"fwd": lambda: call("b200sr_conv3x3_fwd", ptr(x), cin, 0, cin, ptr(wf), cout, B, hw, hw, ptr(out), cout, 0,
                                None, None, 0, ptr(stats), 16, st),
           "dgrad": lambda: call("b200sr_conv3x3_dgrad", ptr(dz), cout, 0, cout, ptr(wf), cin, B, hw, hw, ptr(dx), cin, 0,
                                  None, 0, st),
           "wgrad": lambda: call("b200sr_conv3x3_wgrad", ptr(x), cin, 0, cin, ptr(dz), cout, 0, cout, B, hw, hw, ptr(G), st)}
    line = f"{name:7s} {hw:3d}^2 {cin:4d}->{cout:4d} {flop/1e9:7.1f} GF |"
    for k in which:
        ms = timeit(fns[k])
        tot[k][0] += ms; tot[k][1] += flop
        line += f" {k} {ms*1e3:7.1f} us {flop/ms/1e9:6.0f} TF |"
    print(line, flush=True)
    del x, dz, wf, out, dx, G
for k in which:
    print(f"total {k}: {tot[k][0]:.3f} ms  {tot[k][1]/tot[k][0]/1e9:.0f} TF/s")

# ConvTranspose2d(k2,s2) layers: (name, input hw, Cin, Cout)
if "convT" in sys.argv[1:] or len(sys.argv) == 1:
    for name, hw, cin, cout in [("upconv4", 16, 1024, 512), ("upconv3", 32, 512, 256), ("upconv2", 64, 256, 128),
                                ("upconv1", 128, 128, 64)]:
        x = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
        dup = torch.randn(B, 2 * hw, 2 * hw, 2 * cout, device=dev).to(torch.bfloat16)
        w = (torch.randn(4 * cout * cin, device=dev) * 0.02).to(torch.bfloat16)
        bias = torch.zeros(cout, device=dev)
        dx = torch.empty(B, hw, hw, cin, dtype=torch.bfloat16, device=dev)
        G = torch.zeros(4 * cin * cout, device=dev)
        flop = 2.0 * B * hw * hw * cin * cout * 4
        fns = {"fwd": lambda: call("b200sr_convT2x2_fwd", ptr(x), cin, 0, cin, ptr(w), cout, ptr(bias), B, hw, hw, ptr(dup),
                                    2 * cout, 0, st),
               "dgrad": lambda: call("b200sr_convT2x2_dgrad", ptr(dup), 2 * cout, 0, cout, ptr(w), cin, B, hw, hw, ptr(dx),
                                      cin, 0, st),
               "wgrad": lambda: call("b200sr_convT2x2_wgrad", ptr(dup), 2 * cout, 0, cout, ptr(x), cin, 0, cin, B, hw, hw,
                                      ptr(G), st)}
        line = f"{name:7s} {hw:3d}^2 {cin:4d}->{cout:4d} {flop/1e9:7.1f} GF |"
        mb = (B * hw * hw * cin + B * 4 * hw * hw * cout) * 2 / 1e6
        for k, fn in fns.items():
            ms = timeit(fn)
            line += f" {k} {ms*1e3:7.1f} us {flop/ms/1e9:6.0f} TF {mb/ms/1e3:5.0f} GB/s |"
        print(line, flush=True)
