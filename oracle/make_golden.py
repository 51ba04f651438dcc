"""TEST INFRASTRUCTURE ONLY — pins the oracle against the UNMODIFIED reference and writes tests/golden/.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
  1. imports UNet from /root/reference/src/ModelLoader.py and /root/reference/src/unet_model.py (matplotlib
     stubbed; nothing is copied) and checks both give the state_dict the b200sr mirror gives under the same seed;
  2. checks oracle/unet_oracle.py against the reference module: eval forward, train forward, every gradient,
     updated running statistics, one Adam step (torch.optim.Adam on the reference module);
  3. writes tests/golden/unet_golden.npz with the reference's outputs for the seeded cases in oracle/cases.py.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"

from oracle import cases, fastddpm_oracle, ssim_oracle, unet_oracle  # noqa: E402


def import_reference():
    sys.path.insert(0, REF_SRC)
    for name in ("matplotlib", "matplotlib.pyplot", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "tqdm":
                    m.tqdm = lambda it, **kw: it
                sys.modules[name] = m
    import ModelLoader as ref_loader
    import unet_model as ref_unet
    return ref_loader, ref_unet


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref_loader, ref_unet = import_reference()
    import b200sr

    sd_ref = cases.seeded_state_dict(ref_loader.UNet)
    sd_ref2 = cases.seeded_state_dict(ref_unet.UNet)
    sd_mine = cases.seeded_state_dict(b200sr.UNet)
    assert list(sd_ref) == list(sd_mine) == list(sd_ref2), "state_dict key order differs from the reference"
    for k in sd_ref:
        assert sd_ref[k].shape == sd_mine[k].shape and sd_ref[k].dtype == sd_mine[k].dtype, k
        assert torch.equal(sd_ref[k], sd_mine[k]) and torch.equal(sd_ref[k], sd_ref2[k]), k
    print(f"state_dict: {len(sd_ref)} entries identical (reference ModelLoader.UNet, unet_model.UNet, b200sr.UNet)")

    out = {}
    # ---------------- train case: reference module, MSE (the reference criterion) ----------------
    c = cases.TRAIN_CASE
    x, y = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    model = ref_loader.UNet()
    model.load_state_dict(sd_ref)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    pred = model(x)
    loss = torch.nn.MSELoss()(pred, y)
    opt.zero_grad()
    loss.backward()
    grads_ref = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    # oracle vs reference
    o_loss, o_out, o_grads, o_stats = unet_oracle.loss_and_grads(sd_ref, x, y)
    print(f"train fwd  oracle vs reference rel-L2 {rel(o_out, pred.detach()):.3e}; loss {float(o_loss):.8f} vs {loss.item():.8f}")
    worst = max((rel(o_grads[k], grads_ref[k]), k) for k in grads_ref if grads_ref[k].norm() > 1e-6)
    print(f"train grads oracle vs reference worst rel-L2 {worst[0]:.3e} ({worst[1]})")
    assert rel(o_out, pred.detach()) < 1e-5 and worst[0] < 1e-3
    sd_after = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in o_stats.items():
        assert rel(v, sd_after[k]) < 1e-5, k
    opt.step()
    sd_stepped = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in grads_ref:
        p, _, _ = unet_oracle.adam_update(sd_ref[k], grads_ref[k], torch.zeros_like(sd_ref[k]),
                                          torch.zeros_like(sd_ref[k]), 1)
        assert rel(p - sd_ref[k], sd_stepped[k] - sd_ref[k]) < 1e-4 or (sd_stepped[k] - sd_ref[k]).norm() < 1e-7, k
    print("running stats and one Adam step: oracle == reference")

    out["train_loss"] = np.float64(loss.item())
    out["train_out"] = pred.detach().numpy()
    names = list(grads_ref)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([grads_ref[k].double().norm().item() for k in names])
    out["grad_heads"] = np.stack([np.pad(grads_ref[k].flatten()[:cases.GRAD_HEAD].numpy(),
                                         (0, max(0, cases.GRAD_HEAD - grads_ref[k].numel()))) for k in names])
    stat_names = [k for k in sd_after if k.endswith("running_mean") or k.endswith("running_var")]
    out["stat_names"] = np.array(stat_names)
    out["stats_after"] = np.concatenate([sd_after[k].numpy().ravel() for k in stat_names])
    out["adam_delta_norms"] = np.array([(sd_stepped[k] - sd_ref[k]).double().norm().item() for k in names])
    out["param_sums"] = np.array([sd_ref[k].double().sum().item() for k in names])

    # ---------------- combined loss on the same case (reference module + SSIM oracle; SSIM unpinned) -------
    for mode in ("gaussian", "uniform"):
        model.load_state_dict(sd_ref)
        model.train()
        model.zero_grad()
        pred = model(x)
        closs = ssim_oracle.combined_loss(pred, y, 1.0, 0.005, mode)
        closs.backward()
        out[f"combined_{mode}_loss"] = np.float64(closs.item())
        out[f"combined_{mode}_grad_norms"] = np.array([p.grad.double().norm().item() for _, p in model.named_parameters()])

    # ---------------- eval case: running stats after the train step above -------------------------------
    c = cases.EVAL_CASE
    xe, _ = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    model.load_state_dict(sd_after)
    model.eval()
    with torch.no_grad():
        pe = model(xe)
        oe = unet_oracle.unet_forward(sd_after, xe, training=False)
    print(f"eval fwd   oracle vs reference rel-L2 {rel(oe, pe):.3e}")
    assert rel(oe, pe) < 1e-5
    out["eval_out"] = pe.numpy()

    # ---------------- SSIM oracle vs scipy restatement of skimage (mode uniform) ------------------------
    a, b = x[0, 0].double(), (0.7 * x[0, 0] + 0.3 * y[0, 0]).double()
    s_t = float(ssim_oracle.ssim_map(a[None, None], b[None, None], "uniform").mean())
    s_s = ssim_oracle.ssim_skimage_restatement(a.numpy(), b.numpy())
    print(f"SSIM uniform: torch oracle {s_t:.12f} vs skimage restatement {s_s:.12f}")
    assert abs(s_t - s_s) < 1e-10
    out["ssim_uniform_pair"] = np.float64(s_s)

    path = os.path.join(ROOT, "tests", "golden", "unet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # ---------------- Progressive UNet (3-stage chain) ------------------------------------------
    pout = {}
    sd_p = cases.seeded_state_dict(ref_loader.ProgressiveUNet, seed=3)
    sd_pm = cases.seeded_state_dict(b200sr.ProgressiveUNet, seed=3)
    assert list(sd_p) == list(sd_pm) and all(torch.equal(sd_p[k], sd_pm[k]) for k in sd_p)
    print(f"ProgressiveUNet state_dict: {len(sd_p)} entries identical (reference vs b200sr)")
    c = cases.PROGRESSIVE_CASE
    sl = cases.seeded_slices(c["B"], c["H"], c["W"], c["seed"])
    pm = ref_loader.ProgressiveUNet()
    pm.load_state_dict(sd_p)
    pm.train()
    p1, p2, p3 = pm(sl)
    w = unet_oracle.PROGRESSIVE_LOSS_WEIGHTS
    ploss = sum(wi * torch.nn.functional.mse_loss(p, sl[:, k:k + 1]) for wi, p, k in zip(w, (p1, p2, p3), (1, 2, 3)))
    ploss.backward()
    pg = {k: p.grad.detach().clone() for k, p in pm.named_parameters()}
    o_l, o_p, o_g, _ = unet_oracle.progressive_loss_and_grads(sd_p, sl)
    worst = max(rel(o_g[k], pg[k]) for k in pg if pg[k].norm() > 1e-7)
    print(f"progressive: oracle vs reference outputs {max(rel(a, b.detach()) for a, b in zip(o_p, (p1, p2, p3))):.3e}, "
          f"loss {float(o_l):.8f} vs {ploss.item():.8f}, worst grad rel-L2 {worst:.3e}")
    assert worst < 1e-4 and abs(float(o_l) - ploss.item()) < 1e-6
    pout["keys"] = np.array(list(sd_p))
    pout["loss"] = np.float64(ploss.item())
    pout["p1"], pout["p2"], pout["p3"] = p1.detach().numpy(), p2.detach().numpy(), p3.detach().numpy()
    pnames = list(pg)
    pout["grad_names"] = np.array(pnames)
    pout["grad_norms"] = np.array([pg[k].double().norm().item() for k in pnames])
    pm.eval()
    with torch.no_grad():
        e1, e2, e3 = pm(sl)
    pout["eval_p1"], pout["eval_p2"], pout["eval_p3"] = e1.numpy(), e2.numpy(), e3.numpy()
    ppath = os.path.join(ROOT, "tests", "golden", "progressive_golden.npz")
    np.savez_compressed(ppath, **pout)
    print("wrote", ppath, os.path.getsize(ppath), "bytes")

    # ---------------- DeepCNN residual baseline --------------------------------------------------
    dout = {}
    sd_d = cases.seeded_state_dict(ref_loader.DeepCNN, seed=5)
    sd_dm = cases.seeded_state_dict(b200sr.DeepCNN, seed=5)
    assert list(sd_d) == list(sd_dm) and all(torch.equal(sd_d[k], sd_dm[k]) for k in sd_d)
    print(f"DeepCNN state_dict: {len(sd_d)} entries identical (reference vs b200sr)")
    c = cases.DEEPCNN_CASE
    xd, yd = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    dm = ref_loader.DeepCNN()
    dm.load_state_dict(sd_d)
    dm.train()
    pd_ = dm(xd)
    dloss = torch.nn.functional.mse_loss(pd_, yd)
    dloss.backward()
    dg = {k: p.grad.detach().clone() for k, p in dm.named_parameters()}
    o_l, o_o, o_g, _ = unet_oracle.deepcnn_loss_and_grads(sd_d, xd, yd)
    worst = max(rel(o_g[k], dg[k]) for k in dg if dg[k].norm() > 1e-9)
    print(f"deepcnn: oracle vs reference out {rel(o_o, pd_.detach()):.3e}, loss {float(o_l):.8f} vs {dloss.item():.8f}, "
          f"worst grad rel-L2 {worst:.3e}")
    assert worst < 1e-4 and abs(float(o_l) - dloss.item()) < 1e-6
    dout["keys"] = np.array(list(sd_d))
    dout["loss"] = np.float64(dloss.item())
    dout["train_out"] = pd_.detach().numpy()
    dnames = list(dg)
    dout["grad_names"] = np.array(dnames)
    dout["grad_norms"] = np.array([dg[k].double().norm().item() for k in dnames])
    dm.eval()
    with torch.no_grad():
        dout["eval_out"] = dm(xd).numpy()
    dpath = os.path.join(ROOT, "tests", "golden", "deepcnn_golden.npz")
    np.savez_compressed(dpath, **dout)
    print("wrote", dpath, os.path.getsize(dpath), "bytes")

    # ---------------- UNet-GAN generator: structurally a UNetStage (ModelLoader.py:383-463) --------------------------
    torch.manual_seed(21)
    g_ref = ref_loader.UNetGenerator().state_dict()
    torch.manual_seed(21)
    g_mine = b200sr.UNetGenerator().state_dict()
    assert list(g_ref) == list(g_mine) and all(torch.equal(g_ref[k], g_mine[k]) for k in g_ref)
    print(f"UNetGenerator state_dict: {len(g_ref)} entries identical (reference vs b200sr)")

    # ---------------- Fast-DDPM registry model ---------------------------------------------------
    fout = {}
    sd_f = cases.fastddpm_state_dict(ref_loader.FastDDPM)
    sd_fm = cases.fastddpm_state_dict(b200sr.FastDDPM)
    assert list(sd_f) == list(sd_fm) and all(torch.equal(sd_f[k], sd_fm[k]) for k in sd_f)
    print(f"FastDDPM state_dict: {len(sd_f)} entries identical (reference vs b200sr)")
    ab_o, idx_o = fastddpm_oracle.schedule(10)
    ref_sched = ref_loader.FastNoiseScheduler(10, "cpu")
    mine_sched = b200sr.FastNoiseScheduler(10, "cpu")
    assert torch.equal(ab_o, ref_sched.alpha_bar) and torch.equal(mine_sched.alpha_bar, ref_sched.alpha_bar)
    assert torch.equal(mine_sched.beta, ref_sched.beta) and torch.equal(mine_sched.alpha, ref_sched.alpha)
    for T_ in (4, 5, 20, 50):  # the step selection for other chain lengths (40 % up to t=699, 60 % after)
        r_, m_ = ref_loader.FastNoiseScheduler(T_, "cpu"), b200sr.FastNoiseScheduler(T_, "cpu")
        assert torch.equal(r_.alpha_bar, m_.alpha_bar) and torch.equal(fastddpm_oracle.schedule(T_)[0], r_.alpha_bar), T_
    fout["alpha_bar_T20"] = ref_loader.FastNoiseScheduler(20, "cpu").alpha_bar.numpy()
    c = cases.FASTDDPM_CASE
    cond, target, t, noise = cases.fastddpm_inputs()
    fm = ref_loader.FastDDPM(T=10, device="cpu")
    fm.load_state_dict(sd_f)
    fm.train()
    torch.manual_seed(c["noise_seed"])   # FastDDPM.forward draws torch.randn_like(target) first (:597)
    floss = fm(cond, target, t)
    floss.backward()
    fg = {k: p.grad.detach().clone() for k, p in fm.named_parameters()}
    with torch.no_grad():
        x_t = ref_sched.q_sample(target, t, noise)
        eps_ref = fm.unet(torch.cat([x_t, cond], dim=1), t)
    o_l, o_eps, o_g = fastddpm_oracle.loss_and_grads(sd_f, cond, target, t, noise)
    worst = max(rel(o_g[k], fg[k]) for k in fg if fg[k].norm() > 1e-9)
    print(f"fastddpm: oracle vs reference eps {rel(o_eps, eps_ref):.3e}, loss {float(o_l):.8f} vs {floss.item():.8f}, "
          f"worst grad rel-L2 {worst:.3e}")
    assert worst < 1e-4 and abs(float(o_l) - floss.item()) < 1e-6 and rel(o_eps, eps_ref) < 1e-6
    assert torch.equal(fastddpm_oracle.timestep_embedding(t), ref_loader.sinusoidal_timestep_embedding(t, 256))
    torch.manual_seed(c["noise_seed"] + 1)  # FastDDPM.sample draws torch.randn(B,1,H,W) first (:616)
    s_ref = fm.sample(cond, "cpu")
    torch.manual_seed(c["noise_seed"] + 1)
    x_T = torch.randn(c["B"], 1, c["H"], c["W"])
    s_o = fastddpm_oracle.sample(sd_f, cond, x_T)
    print(f"fastddpm: oracle vs reference 10-step sample {rel(s_o, s_ref):.3e}")
    assert rel(s_o, s_ref) < 1e-5
    fout["keys"] = np.array(list(sd_f))
    fout["alpha_bar"] = ref_sched.alpha_bar.numpy()
    fout["loss"] = np.float64(floss.item())
    fout["eps"] = eps_ref.numpy()
    fnames = list(fg)
    fout["grad_names"] = np.array(fnames)
    fout["grad_norms"] = np.array([fg[k].double().norm().item() for k in fnames])
    for k in fnames:
        fout["grad_head/" + k] = fg[k].reshape(-1)[:cases.GRAD_HEAD].numpy()
    fout["sample"] = s_ref.numpy()
    fpath = os.path.join(ROOT, "tests", "golden", "fastddpm_golden.npz")
    np.savez_compressed(fpath, **fout)
    print("wrote", fpath, os.path.getsize(fpath), "bytes")


if __name__ == "__main__":
    main()
