"""Teacher-forced per-layer parity at the REAL layer shapes of BASELINE configs[0] (B=8, 256x256) and at ragged shapes.

The CPU oracle (pinned bit-for-bit to the reference, oracle/make_golden.py) runs one train-mode forward/backward and
hands out, for each of the 18 Conv3x3+BN+ReLU layers and the 4 ConvTranspose2d layers, its own layer input, raw conv
output, activation and the gradients flowing into them (oracle.unet_oracle.layer_taps). Every CUDA kernel of the path is
then fed THOSE tensors (bf16-rounded, as the path stores them) through the C ABI and must reproduce
  * the oracle's own fp32 result of the layer where the op is linear in its inputs (conv forward, data gradient, weight
    gradient, ConvTranspose forward / gradients): rel-L2 <= 1e-2, the north-star bf16 bound, and
  * a plain PyTorch fp32 evaluation of the same op on the same rounded inputs (TF32 off) for the BatchNorm+ReLU passes,
    whose ReLU mask makes a comparison across differently-rounded inputs ill-posed: rel-L2 <= 1e-2.
This is where the 1e-2 bound is well-posed: no error is carried from layer to layer (the end-to-end gates, calibrated
with the reference's own bf16 run, live in test_gpu_parity_big.py).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 1e-2
SLOTS = 148
WS_FLOATS = 24 * 1024 * 1024


# (B, H, W): the benchmarked layer shapes, plus two inputs whose levels do NOT tile into the 16 x 8 pixel GEMM tile — the
# reference takes any multiple of 16 (src/unet_model.py:56-75). 80x48 -> 40x24 -> 20x12 -> 10x6 -> 5x3 (every level ragged,
# odd sizes, images narrower than a tile); 144x208 -> 72x104 -> 36x52 -> 18x26 -> 9x13 (exact at level 0, ragged below).
SHAPES = [(8, 256, 256), (3, 80, 48), (2, 144, 208)]


@pytest.fixture(scope="module", params=SHAPES, ids=lambda s: "b%d_%dx%d" % s)
def ctx(request):
    import b200sr
    from b200sr import _lib
    from b200sr.engine import _PACK_JOB_DTYPE, _jobs_to_device
    from oracle import cases, ssim_oracle, unet_oracle
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(max(torch.get_num_threads(), 1))
    sd = cases.seeded_state_dict(b200sr.UNet)
    x, y = cases.seeded_batch(*request.param, 4321)
    loss_fn = lambda p, t: ssim_oracle.combined_loss(p, t, 1.0, 0.005, "gaussian")
    loss, out, taps, grads = unet_oracle.layer_taps(sd, x, y, loss_fn)

    class C:
        pass

    c = C()
    c.sd, c.x, c.y, c.taps, c.grads, c.lib = sd, x, y, taps, grads, _lib
    c.ws = torch.empty(WS_FLOATS, dtype=torch.float32, device="cuda")
    c.counters = torch.zeros(64, dtype=torch.int32, device="cuda")

    def pack(src, kind, cout, cin):
        dst = torch.zeros(src.numel(), dtype=torch.bfloat16, device="cuda")
        job = np.zeros(1, dtype=_PACK_JOB_DTYPE)
        job[0] = (src.data_ptr(), dst.data_ptr(), kind, cout, cin, 0, src.numel())
        jobs = _jobs_to_device(job, "cuda")
        _lib.call("b200sr_pack_jobs", jobs.data_ptr(), 1, _lib.current_stream_ptr())
        torch.cuda.synchronize()
        return dst

    c.pack = pack
    return c


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(t):
    return t.cuda().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def bf(t):
    return t.to(torch.bfloat16).float()


CONV_LAYERS = [f"{b}.conv.{i}" for b in ("enc1", "enc2", "enc3", "enc4", "bottleneck", "dec4", "dec3", "dec2", "dec1")
               for i in (0, 3)]


@pytest.mark.parametrize("name", CONV_LAYERS)
def test_conv_bn_relu_layer(ctx, name):
    from b200sr._lib import call, ptr
    st = ctx.lib.current_stream_ptr()
    tap = ctx.taps[name]
    blk, idx = name.rsplit(".conv.", 1)
    bn_name = f"{blk}.conv.{int(idx) + 1}"
    w = ctx.sd[f"{name}.weight"].cuda()
    bias = ctx.sd[f"{name}.bias"].cuda()
    gamma, beta = ctx.sd[f"{bn_name}.weight"].cuda(), ctx.sd[f"{bn_name}.bias"].cuda()
    Cout, Cin = w.shape[0], w.shape[1]
    a_in, z_ref, act_ref, dz_ref, dact_ref = (tap[k] for k in ("a_in", "z", "act", "dz", "dact"))
    B, _, H, W = a_in.shape
    npix = B * H * W
    res = {}

    # ---- forward: conv (+ per-CTA BatchNorm statistics) -> finalize -> BN-apply + ReLU -------------------------------
    zb = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
    first = Cin == 2
    slots = 2 * SLOTS if first else SLOTS
    stats = torch.full((slots, 2, Cout), 7.0, device="cuda")
    if first:
        xin = a_in.cuda().contiguous()
        call("b200sr_conv1_fwd", ptr(xin), ptr(w), None, None, 0, ptr(zb), ptr(stats), slots, B, H, W, st)
    else:
        ab = nhwc(a_in)
        wp = ctx.pack(w, 0, Cout, Cin)
        call("b200sr_conv3x3_fwd", ptr(ab), Cin, 0, Cin, ptr(wp), Cout, B, H, W, ptr(zb), Cout, 0, None, None, 0,
             ptr(stats), slots, st)
    ws = torch.zeros(4, Cout, device="cuda")
    rm, rv = torch.zeros(Cout, device="cuda"), torch.ones(Cout, device="cuda")
    call("b200sr_bn_finalize", ptr(stats), slots, Cout, float(npix), ptr(gamma), ptr(beta), ptr(bias), 1e-5, 0.1,
         ptr(ws[0]), ptr(ws[1]), ptr(ws[2]), ptr(ws[3]), ptr(rm), ptr(rv), None, st)
    actb = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
    call("b200sr_bnrelu_apply", ptr(zb), Cout, ptr(ws[0]), ptr(ws[1]), ptr(actb), Cout, 0, None, B, H, W, st)
    torch.cuda.synchronize()
    z_cuda = nchw(zb) + bias[None, :, None, None]   # the path stores z without the bias BatchNorm cancels
    res["z"] = rel(z_cuda.cpu(), z_ref)
    res["act"] = rel(nchw(actb).cpu(), act_ref)
    # running statistics: mean incl. conv bias, unbiased variance (momentum 0.1 from 0 / 1)
    res["running_mean"] = rel(rm.cpu(), 0.1 * z_ref.mean(dim=(0, 2, 3)))
    res["running_var"] = rel(rv.cpu(), 0.9 + 0.1 * z_ref.var(dim=(0, 2, 3), unbiased=True))

    # ---- BatchNorm + ReLU backward on the oracle's dact and the stored z: vs torch fp32 on the SAME inputs ------------
    z_in = nchw(zb).requires_grad_(True)   # the bf16 z the kernels stored
    g_t, b_t = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    dact = bf(dact_ref.cuda())
    torch.relu(F.batch_norm(z_in, None, None, g_t, b_t, True, 0.1, 1e-5)).backward(dact)
    dactb = nhwc(dact_ref)
    sums = torch.full((2, Cout), 7.0, device="cuda")
    dgb = torch.full((2, Cout), 7.0, device="cuda")
    dzb = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
    call("b200sr_bn_bwd_reduce_det", ptr(dactb), Cout, 0, ptr(zb), Cout, ptr(ws[0]), ptr(ws[1]), ptr(ws[2]), ptr(ws[3]),
         ptr(sums), ptr(ctx.ws), ctx.ws.numel(), ptr(ctx.counters), None, npix, st)
    call("b200sr_bn_bwd_apply_fused", ptr(dactb), Cout, 0, ptr(zb), Cout, ptr(ws[0]), ptr(ws[1]), ptr(ws[2]), ptr(ws[3]),
         ptr(sums), 1, float(npix), ptr(dgb[0]), ptr(dgb[1]), ptr(dzb), npix, st)
    torch.cuda.synchronize()
    res["bn_bwd_dz"] = rel(nchw(dzb), z_in.grad)
    res["bn_bwd_dgamma"] = rel(dgb[0], g_t.grad)
    res["bn_bwd_dbeta"] = rel(dgb[1], b_t.grad)
    res["bn_bwd_dz_vs_oracle(info)"] = rel(nchw(dzb).cpu(), dz_ref)

    # ---- weight gradient from the oracle's input and dz: vs the oracle's own fp32 gradient -----------------------------
    dzo = nhwc(dz_ref)
    dw = torch.full((Cout, Cin, 3, 3), 7.0, device="cuda")
    if first:
        call("b200sr_conv1_wgrad_det", ptr(xin), ptr(dzo), ptr(dw), B, H, W, ptr(ctx.ws), ctx.ws.numel(), st)
    else:
        call("b200sr_conv3x3_wgrad_det", ptr(ab), Cin, 0, Cin, ptr(dzo), Cout, 0, Cout, B, H, W, ptr(dw), Cin, 0,
             ptr(ctx.ws), ctx.ws.numel(), st)
    torch.cuda.synchronize()
    res["wgrad"] = rel(dw.cpu(), ctx.grads[f"{name}.weight"])

    # ---- data gradient from the oracle's dz: vs fp32 conv_transpose2d of the unrounded dz ------------------------------
    if not first:
        wpd = ctx.pack(w, 1, Cout, Cin)
        dxb = torch.zeros(B, H, W, Cin, dtype=torch.bfloat16, device="cuda")
        call("b200sr_conv3x3_dgrad", ptr(dzo), Cout, 0, Cout, ptr(wpd), Cin, B, H, W, ptr(dxb), Cin, 0, None, 0, st)
        torch.cuda.synchronize()
        res["dgrad"] = rel(nchw(dxb), F.conv_transpose2d(dz_ref.cuda(), w, padding=1))
    else:
        dx = torch.zeros(B, 2, H, W, device="cuda")
        call("b200sr_conv1_dgrad", ptr(dzo), ptr(w), ptr(dx), B, H, W, st)
        torch.cuda.synchronize()
        res["dgrad"] = rel(dx, F.conv_transpose2d(dz_ref.cuda(), w, padding=1))

    bad = {k: v for k, v in res.items() if not k.endswith("(info)") and not v <= TOL}
    assert not bad, f"{name} {tuple(a_in.shape)}->{Cout}: {bad}; all: {res}"


@pytest.mark.parametrize("k", [4, 3, 2, 1])
def test_conv_transpose_layer(ctx, k):
    from b200sr._lib import call, ptr
    st = ctx.lib.current_stream_ptr()
    name = f"upconv{k}"
    tap = ctx.taps[name]
    w, bias = ctx.sd[f"{name}.weight"].cuda(), ctx.sd[f"{name}.bias"].cuda()
    Cin, Cout = w.shape[0], w.shape[1]
    a_in, out_ref, dout_ref = tap["a_in"], tap["out"], tap["dout"]
    B, _, H, W = a_in.shape
    ab = nhwc(a_in)
    res = {}
    # forward into the [0, Cout) slot of a concat buffer
    wpf = ctx.pack(w, 2, Cout, Cin)
    cat = torch.full((B, 2 * H, 2 * W, 2 * Cout), 7.0, dtype=torch.bfloat16, device="cuda")
    call("b200sr_convT2x2_fwd", ptr(ab), Cin, 0, Cin, ptr(wpf), Cout, ptr(bias), B, H, W, ptr(cat), 2 * Cout, 0, st)
    torch.cuda.synchronize()
    res["fwd"] = rel(nchw(cat[..., :Cout]).cpu(), out_ref)
    res["slot_untouched"] = float((cat[..., Cout:].float() - 7.0).abs().max())
    # data gradient
    dob = torch.full((B, 2 * H, 2 * W, 2 * Cout), 7.0, dtype=torch.bfloat16, device="cuda")
    dob[..., :Cout] = nhwc(dout_ref)
    wpd = ctx.pack(w, 3, Cout, Cin)
    dxb = torch.zeros(B, H, W, Cin, dtype=torch.bfloat16, device="cuda")
    call("b200sr_convT2x2_dgrad", ptr(dob), 2 * Cout, 0, Cout, ptr(wpd), Cin, B, H, W, ptr(dxb), Cin, 0, st)
    torch.cuda.synchronize()
    res["dgrad"] = rel(nchw(dxb), F.conv2d(dout_ref.cuda(), w, stride=2))  # dgrad of ConvT(k2,s2) = conv(k2,s2)
    # weight gradient
    dw = torch.full((Cin, Cout, 2, 2), 7.0, device="cuda")
    call("b200sr_convT2x2_wgrad_det", ptr(dob), 2 * Cout, 0, Cout, ptr(ab), Cin, 0, Cin, B, H, W, ptr(dw), ptr(ctx.ws),
         ctx.ws.numel(), st)
    torch.cuda.synchronize()
    res["wgrad"] = rel(dw.cpu(), ctx.grads[f"{name}.weight"])
    res["bias_grad"] = rel(dout_ref.sum(dim=(0, 2, 3)), ctx.grads[f"{name}.bias"])  # oracle self-consistency
    bad = {kk: v for kk, v in res.items() if not v <= (0.0 if kk == "slot_untouched" else TOL)}
    assert not bad, f"{name}: {bad}; all: {res}"


@pytest.mark.parametrize("lvl", [0, 1, 2, 3])
def test_maxpool_level(ctx, lvl):
    """MaxPool2d(2,2) forward / backward(+skip add) on the oracle's encoder activations (bf16-rounded): exact."""
    from b200sr._lib import call, ptr
    st = ctx.lib.current_stream_ptr()
    act_ref = ctx.taps[f"enc{lvl + 1}.conv.3"]["act"]
    B, C, H, W = act_ref.shape
    a = bf(act_ref.cuda()).requires_grad_(True)
    pooled_ref = F.max_pool2d(a, 2, 2)
    g = torch.Generator(device="cpu").manual_seed(100 + lvl)
    dpool = bf(torch.randn(pooled_ref.shape, generator=g).cuda())
    dskip = bf(torch.randn(a.shape, generator=g).cuda())
    pooled_ref.backward(dpool)
    dy_ref = bf(a.grad + dskip)
    ab = nhwc(a.detach())
    pb = torch.zeros(B, H // 2, W // 2, C, dtype=torch.bfloat16, device="cuda")
    call("b200sr_maxpool2x2_fwd", ptr(ab), C, 0, C, ptr(pb), B, H, W, st)
    dyb = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    dpb, dsb = nhwc(dpool), nhwc(dskip)   # named: a temporary would be freed (and its memory reused) before the launch
    call("b200sr_maxpool2x2_bwd", ptr(ab), C, 0, ptr(dpb), ptr(dsb), C, 0, C, ptr(dyb), B, H, W, st)
    torch.cuda.synchronize()
    assert torch.equal(nchw(pb), pooled_ref.detach())
    assert float((nchw(dyb) - dy_ref).abs().max()) == 0.0


def test_head_and_loss(ctx):
    """final 1x1 conv on the oracle's last activation, then the fused MSE+SSIM loss value / gradient on the oracle's output."""
    import b200sr
    from b200sr._lib import call, ptr
    from oracle import ssim_oracle, unet_oracle
    st = ctx.lib.current_stream_ptr()
    act = ctx.taps["dec1.conv.3"]["act"]
    B, C, H, W = act.shape
    w, b = ctx.sd["final_conv.weight"].cuda(), ctx.sd["final_conv.bias"].cuda()
    out = torch.zeros(B, 1, H, W, device="cuda")
    actb = nhwc(act)
    call("b200sr_head_fwd", ptr(actb), ptr(w), ptr(b), ptr(out), B * H * W, st)
    out_ref = F.conv2d(act, ctx.sd["final_conv.weight"], ctx.sd["final_conv.bias"])
    assert rel(out.cpu(), out_ref) <= TOL
    p = out_ref.clone().requires_grad_(True)
    loss_ref = ssim_oracle.combined_loss(p, ctx.y, 1.0, 0.005, "gaussian")
    loss_ref.backward()
    crit = b200sr.CombinedLoss(1.0, 0.005, "gaussian")
    loss, grad = crit.value_and_grad(out_ref.cuda(), ctx.y.cuda())
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)) <= 1e-5
    assert rel(grad.cpu(), p.grad) <= 1e-4
