"""GPU parity of the Progressive UNet chain against the CPU oracle / golden fixtures. Gates as in
test_gpu_parity_big.py: loss 1e-3 relative; outputs and every gradient gated PER TENSOR by the unmodified reference's own
bf16-autocast deviation on this case (tests/golden/bf16_calibration.json `progressive_small`, oracle/make_calibration.py):
rel-L2 <= max(1e-2, 1.5 x reference), cosine >= 0.9 (0.99 for the heads / dec1) unless the reference itself is below."""
import json
import os

import numpy as np
import pytest
import torch

import b200sr
from oracle import cases, unet_oracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "progressive_golden.npz"))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bf16_calibration.json")) as _f:
    CAL = json.load(_f)["progressive_small"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _model(sd):
    m = b200sr.ProgressiveUNet()
    m.load_state_dict(sd)
    return m.cuda()


def test_eval_forward_matches_golden():
    sd = cases.seeded_state_dict(b200sr.ProgressiveUNet, seed=3)
    c = cases.PROGRESSIVE_CASE
    sl = cases.seeded_slices(c["B"], c["H"], c["W"], c["seed"])
    _, _, _, stats = unet_oracle.progressive_loss_and_grads(sd, sl)
    sd.update(stats)  # the golden eval pass ran after one train-mode forward (updated running statistics)
    m = _model(sd).eval()
    with torch.no_grad():
        outs = m(sl.cuda())
    for o, k in zip(outs, ("eval_p1", "eval_p2", "eval_p3")):
        assert rel(o.cpu(), torch.from_numpy(GOLD[k])) < 1e-2, k   # eval mode: the north-star bf16 bound


def test_train_step_matches_oracle():
    sd = cases.seeded_state_dict(b200sr.ProgressiveUNet, seed=3)
    c = cases.PROGRESSIVE_CASE
    sl = cases.seeded_slices(c["B"], c["H"], c["W"], c["seed"])
    o_loss, o_out, o_grads, _ = unet_oracle.progressive_loss_and_grads(sd, sl)
    assert abs(float(o_loss) - float(GOLD["loss"])) < 1e-6
    m = _model(sd).train()
    # autograd path: the three stages chained by torch (cat / slicing), gradients flow into stage 1 through p2
    x = sl.cuda()
    p1, p2, p3 = m(x)
    crit = [b200sr.CombinedLoss(w, 0.0) for w in unet_oracle.PROGRESSIVE_LOSS_WEIGHTS]
    loss = crit[0](p1, x[:, 1:2].contiguous()) + crit[1](p2, x[:, 2:3].contiguous()) + crit[2](p3, x[:, 3:4].contiguous())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(o_loss)) / float(o_loss) < 1e-3
    assert abs(float(o_loss) - CAL["loss_fp32"]) / CAL["loss_fp32"] < 1e-5   # same case as the calibration run
    for got, ref, r_cal in zip((p1, p2, p3), o_out, CAL["out"]):
        assert rel(got.detach().cpu(), ref) <= max(1e-2, 1.5 * r_cal), (rel(got.detach().cpu(), ref), r_cal)
    for name, p in m.named_parameters():
        ref = o_grads[name]
        if ref.norm() < 1e-7:
            continue
        r, cs = rel(p.grad.cpu(), ref), cos(p.grad.cpu(), ref)
        r_cal, c_cal = CAL["grads"][name]
        shallow = name.split(".", 1)[1].startswith(("final", "dec1"))
        assert r <= max(1e-2, 1.5 * r_cal), (name, r, r_cal)
        assert cs >= min(0.99 if shallow else 0.9, 1.0 - 1.5 * (1.0 - c_cal)), (name, cs, c_cal)


def test_trainer_matches_autograd_and_learns(tmp_path):
    sd = cases.seeded_state_dict(b200sr.ProgressiveUNet, seed=3)
    sl = cases.seeded_slices(2, 128, 256, 77).cuda()
    m = _model(sd)
    tr = b200sr.ProgressiveUNetTrainer(m, device="cuda", learning_rate=5e-4, model_save_dir=str(tmp_path), verbose=False)
    losses = [float(tr.train_step(sl)) for _ in range(6)]
    assert losses[-1] < losses[0], losses
    tr.save_checkpoint(1, losses[-1], is_best=True)
    ck = torch.load(tmp_path / "progressive_unet_best.pt", map_location="cpu")
    assert len(ck["model_state_dict"]) == 354
