"""Whole-network parity of the CUDA path against the CPU oracle and the committed golden fixtures."""
from __future__ import annotations

import os

import numpy as np
import torch

import b200sr
from oracle import cases, ssim_oracle, unet_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "unet_golden.npz")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def build_model(sd):
    m = b200sr.UNet()
    m.load_state_dict(sd)
    return m.cuda()


def train_case(loss="mse", mode="gaussian"):
    """Returns dict of errors of one train-mode forward/backward vs the oracle (and golden where stored)."""
    gold = np.load(GOLDEN)
    sd = cases.seeded_state_dict(b200sr.UNet)
    c = cases.TRAIN_CASE
    x, y = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    if loss == "mse":
        loss_fn, crit = None, b200sr.CombinedLoss(1.0, 0.0, mode)
    else:
        loss_fn = lambda p, t: ssim_oracle.combined_loss(p, t, 1.0, 0.005, mode)
        crit = b200sr.CombinedLoss(1.0, 0.005, mode)
    o_loss, o_out, o_grads, o_stats = unet_oracle.loss_and_grads(sd, x, y, loss_fn)
    res = {}
    if loss == "mse":
        res["oracle_vs_golden_out"] = rel(o_out, torch.from_numpy(gold["train_out"]))
        res["oracle_vs_golden_loss"] = abs(float(o_loss) - float(gold["train_loss"])) / float(gold["train_loss"])

    model = build_model(sd)
    model.train()
    eng = model._get_engine()
    out = eng.forward_train(x.cuda())
    lval, dout = crit.value_and_grad(out, y.cuda())
    eng.backward(dout)
    torch.cuda.synchronize()
    res["out"] = rel(out.cpu(), o_out)
    res["loss"] = abs(float(lval) - float(o_loss)) / abs(float(o_loss))
    names = unet_oracle.param_names(sd)
    grads = {n: g.detach().float().cpu().clone() for n, g in zip(names, eng.grad_views)}
    per = {}
    for n in names:
        ref = o_grads[n]
        if n.endswith("conv.0.bias") or n.endswith("conv.3.bias"):
            per[n] = ("abs", float(grads[n].abs().max()))  # BN cancels these: true gradient is 0
        else:
            per[n] = ("rel", rel(grads[n], ref), cos(grads[n], ref))
    res["grads"] = per
    # running statistics after the step
    msd = model.state_dict()
    res["running_stats"] = max(rel(msd[k].cpu(), v) for k, v in o_stats.items())
    res["num_batches_tracked"] = int(msd["enc1.conv.1.num_batches_tracked"])
    return res


def eval_case():
    gold = np.load(GOLDEN)
    sd = cases.seeded_state_dict(b200sr.UNet)
    c = cases.TRAIN_CASE
    x, y = cases.seeded_batch(c["B"], c["H"], c["W"], c["seed"])
    _, _, _, new_stats = unet_oracle.loss_and_grads(sd, x, y)
    sd.update(new_stats)
    ce = cases.EVAL_CASE
    xe, _ = cases.seeded_batch(ce["B"], ce["H"], ce["W"], ce["seed"])
    ref = unet_oracle.unet_forward(sd, xe, training=False)
    model = build_model(sd).eval()
    with torch.no_grad():
        out = model(xe.cuda())
    torch.cuda.synchronize()
    return {"out_vs_oracle": rel(out.cpu(), ref), "out_vs_golden": rel(out.cpu(), torch.from_numpy(gold["eval_out"])),
            "oracle_vs_golden": rel(ref, torch.from_numpy(gold["eval_out"]))}


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    e = eval_case()
    print("EVAL", e)
    for loss in ("mse", "combined"):
        r = train_case(loss)
        g = r.pop("grads")
        print("TRAIN", loss, r)
        for n, v in g.items():
            print(f"   {n:28s} {v}")
