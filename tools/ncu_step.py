"""One profiled UNet train step at the benchmarked configuration (B=32, 256x256, combined loss), for ncu:
3 warm-up steps, then cudaProfilerStart / one step / cudaProfilerStop.
   python tools/ncu_step.py                  (plain run: must exit 0 before it is profiled)
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/ncu_step.py
   ncu --profile-from-start off --set full --clock-control none -k regex:<kernels> -o X python tools/ncu_step.py
Options: --eval (profile one B=8 eval forward instead), --fp32 (eval in the fp32-accuracy mode), --batch N."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200sr  # noqa: E402
from oracle import cases  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--eval", action="store_true")
ap.add_argument("--fp32", action="store_true")
args = ap.parse_args()
os.environ.setdefault("B200SR_NO_EVAL_GRAPH", "1")  # graphs hide the kernels from a per-launch profile

dev = torch.device("cuda", 0)
model = b200sr.UNet()
model.load_state_dict(cases.seeded_state_dict(b200sr.UNet))
gen = b200sr.SyntheticTripletGenerator(args.batch, 256, 256, device=dev, seed=1234, rank=0)
batches = [gen.next() for _ in range(4)]
if args.eval:
    model = model.cuda().eval().set_eval_precision("fp32" if args.fp32 else "bf16")
    x = batches[0][0][:8].contiguous()
    with torch.no_grad():
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        model(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
else:
    tr = b200sr.UNetTrainer(model, device=dev, loss="combined", ssim_weight=0.005, model_save_dir="/tmp/b200sr_ncu",
                            verbose=False)
    for i in range(3):
        tr.train_step(*batches[i])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = tr.train_step(*batches[3])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", float(loss))
print("ncu_step done")
