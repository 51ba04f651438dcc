"""GPU parity of the Fast-DDPM row against the CPU oracle / golden fixtures (from the unmodified
reference). Tolerances: bf16 tensor-core compute -> rel-L2 1e-2 on eps and on the 10-step sample, 1e-3 relative on the
loss (north star), gradients 2e-2 rel-L2 / cosine; the fp32 CUDA-core pieces (time MLP, time-embedding fold, q_sample,
DDIM update, grad clip) are checked at 1e-5..1e-4."""
import os

import numpy as np
import pytest
import torch

import b200sr
from b200sr import _lib
from b200sr._lib import call, ptr
from oracle import cases, fastddpm_oracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastddpm_golden.npz"))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _model(sd):
    m = b200sr.FastDDPM(T=10, device="cpu")
    m.load_state_dict(sd)
    m = m.cuda()
    m.scheduler.to("cuda")
    return m


def test_time_mlp_and_bias_table_fp32():
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    m = _model(sd)
    eng = m.unet._get_engine()
    eng.ensure_ready(torch.device("cuda", 0))
    B = 5
    t = torch.tensor([0, 3, 9, 4, 7], device="cuda")
    emb, hid, e = (torch.empty(B, 256, device="cuda") for _ in range(3))
    l1, l2 = m.unet.time_mlp[0], m.unet.time_mlp[2]
    st = _lib.current_stream_ptr()
    call("b200sr_fd_time_mlp_fwd", ptr(t), ptr(l1.weight), ptr(l1.bias), ptr(l2.weight), ptr(l2.bias), ptr(emb), ptr(hid),
         ptr(e), B, st)
    ref_emb = fastddpm_oracle.timestep_embedding(t.cpu())
    ref_e = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(ref_emb, sd["unet.time_mlp.0.weight"],
                                                                             sd["unet.time_mlp.0.bias"])),
                                       sd["unet.time_mlp.2.weight"], sd["unet.time_mlp.2.bias"])
    assert rel(emb.cpu(), ref_emb) < 1e-5 and rel(e.cpu(), ref_e) < 1e-5
    tb = torch.empty(B, 9, 64, device="cuda")
    w = m.unet.inc.block[0]
    call("b200sr_fd_time_bias", ptr(e), ptr(w.weight), ptr(w.bias), ptr(tb), B, st)
    # reference: conv of the tiled embedding on a 3x3 image hits all 9 border classes once
    tiled = ref_e[:, :, None, None].repeat(1, 1, 3, 3)
    ref_tb = torch.nn.functional.conv2d(tiled, sd["unet.inc.block.0.weight"][:, 3:], sd["unet.inc.block.0.bias"], padding=1)
    assert rel(tb.cpu().permute(0, 2, 1).reshape(B, 64, 3, 3), ref_tb) < 1e-5


def test_q_sample_and_ddim_update_fp32():
    s = b200sr.FastNoiseScheduler(10, "cuda")
    g = torch.Generator().manual_seed(1)
    x0, noise = torch.randn(3, 1, 32, 48, generator=g), torch.randn(3, 1, 32, 48, generator=g)
    t = torch.tensor([0, 4, 9])
    out = s.q_sample(x0.cuda(), t.cuda(), noise.cuda())
    assert rel(out.cpu(), fastddpm_oracle.q_sample(x0, t, noise)) < 1e-6
    ab = fastddpm_oracle.schedule(10)[0]
    x, eps = x0.clone().cuda(), noise.cuda()
    call("b200sr_fd_ddim_update", ptr(x), ptr(eps), float(ab[5]), float(ab[4]), 0, x.numel(), _lib.current_stream_ptr())
    xr = (x0 - torch.sqrt(1 - ab[5]) * noise) / torch.sqrt(ab[5])
    xr = torch.sqrt(ab[4]) * xr + torch.sqrt(1 - ab[4]) * noise
    assert rel(x.cpu(), xr) < 1e-6


def test_train_forward_backward_matches_oracle():
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    cond, target, t, noise = cases.fastddpm_inputs()
    o_loss, o_eps, o_grads = fastddpm_oracle.loss_and_grads(sd, cond, target, t, noise)
    assert abs(float(o_loss) - float(GOLD["loss"])) / float(GOLD["loss"]) < 1e-6
    m = _model(sd).train()
    eng = m.unet._get_engine()
    loss, dout = m.loss_and_grad(cond.cuda(), target.cuda(), t.cuda(), noise=noise.cuda())
    eps = eng.forward(target.cuda(), cond.cuda(), t.cuda(), noise=noise.cuda(), coef=m._coef(t.cuda(), target.cuda().device),
                      keep=True)
    eng.backward(dout)
    torch.cuda.synchronize()
    assert rel(eps.cpu(), torch.from_numpy(GOLD["eps"])) < 1e-2
    assert abs(float(loss) - float(GOLD["loss"])) / float(GOLD["loss"]) < 1e-3
    worst = 0.0
    for (name, _), g in zip(m.unet.named_parameters(), eng.grad_views):
        ref = o_grads["unet." + name]
        r, cs = rel(g.cpu(), ref), cos(g.cpu(), ref)
        worst = max(worst, r)
        assert cs > 0.999 and r < 3e-2, (name, r, cs)
    print("fastddpm worst gradient rel-L2", worst)


def test_forward_loss_is_differentiable_like_the_reference():
    """Reference FastDDPM.forward returns a differentiable F.mse_loss and its training loops call loss.backward()
    (ModelLoader.py:595-602): the drop-in must fill p.grad the same way (autograd node around the engine)."""
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    cond, target, t, noise = cases.fastddpm_inputs()
    o_loss, _, o_grads = fastddpm_oracle.loss_and_grads(sd, cond, target, t, noise)
    m = _model(sd).train()
    loss = m(cond.cuda(), target.cuda(), t.cuda(), noise=noise.cuda())
    assert loss.requires_grad
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(o_loss)) / float(o_loss) < 1e-3
    for name, p in m.unet.named_parameters():
        assert p.grad is not None, name
        ref = o_grads["unet." + name]
        assert cos(p.grad.cpu(), ref) > 0.999 and rel(p.grad.cpu(), ref) < 3e-2, name
    with torch.no_grad():
        assert not m(cond.cuda(), target.cuda(), t.cuda(), noise=noise.cuda()).requires_grad


def test_unet2d_module_call_and_autograd():
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    cond, target, t, noise = cases.fastddpm_inputs()
    m = _model(sd)
    x_t = fastddpm_oracle.q_sample(target, t, noise)
    x = torch.cat([x_t, cond], 1).cuda()
    eps = m.unet(x, t.cuda())
    loss = torch.nn.functional.mse_loss(eps, noise.cuda())
    loss.backward()
    assert rel(eps.detach().cpu(), torch.from_numpy(GOLD["eps"])) < 1e-2
    g = m.unet.outc.weight.grad
    assert g is not None and abs(float(g.double().norm()) - float(GOLD["grad_norms"][list(GOLD["grad_names"]).index(
        "unet.outc.weight")])) / float(g.double().norm()) < 2e-2
    with torch.no_grad():
        eps2 = m.unet(x, t.cuda())
    assert torch.equal(eps2, eps.detach())


def test_sample_matches_golden():
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    cond, _, _, _ = cases.fastddpm_inputs()
    c = cases.FASTDDPM_CASE
    torch.manual_seed(c["noise_seed"] + 1)
    x_T = torch.randn(c["B"], 1, c["H"], c["W"])
    m = _model(sd).eval()
    out = m.sample(cond.cuda(), "cuda", noise=x_T.cuda())
    assert out.shape == (c["B"], 1, c["H"], c["W"]) and float(out.abs().max()) <= 1.0
    assert rel(out.cpu(), torch.from_numpy(GOLD["sample"])) < 1e-2
    # from the third call per shape the whole 10-step chain is replayed from one CUDA graph: same bits as the eager run
    again = [m.sample(cond.cuda(), "cuda", noise=x_T.cuda()) for _ in range(3)]
    assert all(torch.equal(out, o) for o in again)
    out2 = m.sample(cond.cuda(), "cuda")  # draws its own x_T like the reference
    assert out2.shape == out.shape and torch.isfinite(out2).all()


def test_grad_clip_and_trainer_step_match_oracle(tmp_path):
    sd = cases.fastddpm_state_dict(b200sr.FastDDPM)
    cond, target, t, noise = cases.fastddpm_inputs()
    _, _, o_grads = fastddpm_oracle.loss_and_grads(sd, cond, target, t, noise)
    new_sd = fastddpm_oracle.clip_and_adam(sd, o_grads, {}, 1, lr=2e-4, max_norm=1.0)
    m = _model(sd)
    tr = b200sr.FastDDPMTrainer(m, device="cuda", learning_rate=2e-4, grad_clip=1.0, model_save_dir=str(tmp_path),
                                verbose=False)
    tr.train_step(cond.cuda(), target.cuda(), t=t.cuda(), noise=noise.cuda())
    torch.cuda.synchronize()
    # first Adam step moves every weight by ~lr*sign(g): compare the update direction where the gradient is not tiny
    for k, v in m.state_dict().items():
        upd, ref = v.cpu() - sd[k], new_sd[k] - sd[k]
        # Adam's first step is lr*sign(g): elements whose gradient is within bf16 noise of zero may flip, so the sign is
        # compared where the gradient is at least 5 % of the tensor's largest, and the whole update by its cosine
        big = o_grads[k].abs() > 5e-2 * o_grads[k].abs().max()
        agree = (torch.sign(upd[big]) == torch.sign(ref[big])).float().mean()
        assert agree > 0.97, (k, float(agree))
        assert cos(upd, ref) > 0.9, (k, cos(upd, ref))
        assert float(upd.abs().max()) <= 2e-4 * 1.01
    # clip kernel on its own: fp32 exactness
    g = torch.randn(100_003, device="cuda") * 3
    ref = g.clone().cpu()
    ssq = torch.zeros(1, dtype=torch.float64, device="cuda")
    call("b200sr_grad_clip", ptr(g), g.numel(), ptr(ssq), 1.0, 0.5, _lib.current_stream_ptr())
    coef = min(1.0, 1.0 / (0.5 * float(ref.double().norm()) + 1e-6))
    assert rel(g.cpu(), ref * coef) < 1e-5


def test_trainer_learns(tmp_path):
    torch.manual_seed(0)
    m = b200sr.FastDDPM(T=10, device="cuda")
    tr = b200sr.FastDDPMTrainer(m, device="cuda", learning_rate=2e-4, model_save_dir=str(tmp_path), verbose=False)
    gen = b200sr.SyntheticTripletGenerator(4, 64, 64, device="cuda", seed=3)
    x, y = gen.next()
    g = torch.Generator(device="cuda").manual_seed(5)
    t = torch.randint(0, 10, (4,), device="cuda", generator=g)
    noise = torch.randn(4, 1, 64, 64, device="cuda", generator=g)
    losses = [float(tr.train_step(x, y, t=t, noise=noise)) for _ in range(40)]
    # noise prediction from a random init moves slowly at the notebook's learning rate (measured: 1.019 -> 0.962 in 30
    # steps, monotone); the gate is a steady decrease, not a large one
    assert losses[-1] < 0.97 * losses[0] and losses[-1] < losses[10] < losses[0], losses


def test_unet_generator_runs_on_the_unet_kernels():
    torch.manual_seed(2)
    g = b200sr.UNetGenerator().cuda().eval()
    x = torch.randn(1, 2, 128, 256, device="cuda")
    with torch.no_grad():
        y = g(x)
    assert y.shape == (1, 1, 128, 256) and torch.isfinite(y).all()
