"""Train-step throughput of the workload variants beyond the headline (B=32/GPU, 256x256, one GPU):
   combined_perceptual (MSE + VGG16 perceptual + SSIM: the full BASELINE configs[2]) and the Progressive UNet chain."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200sr
from oracle import cases

dev = "cuda"


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(which=("perceptual", "progressive", "deepcnn", "fastddpm"), B=32):
    """Returns {variant: {...}} measured on the current CUDA device (single process, no collectives)."""
    out = {}
    if "perceptual" in which:
        model = b200sr.UNet()
        model.load_state_dict(cases.seeded_state_dict(b200sr.UNet))
        tr = b200sr.UNetTrainer(model, device=dev, loss="combined_perceptual", model_save_dir="/tmp/b200sr_v", verbose=False)
        gen = b200sr.SyntheticTripletGenerator(B, 256, 256, device=dev, seed=1)
        x, y = gen.next()
        ms = timed(lambda: tr.train_step(x, y))
        out["unet_combined_perceptual"] = {"ms_per_step": ms, "triplets_per_s": B / ms * 1e3}
        del tr, model
        torch.cuda.empty_cache()
    if "progressive" in which:
        pm = b200sr.ProgressiveUNet()
        ptr_ = b200sr.ProgressiveUNetTrainer(pm, device=dev, model_save_dir="/tmp/b200sr_v", verbose=False)
        sl = cases.seeded_slices(B, 256, 256, 5).to(dev)
        ms = timed(lambda: ptr_.train_step(sl), steps=6, warmup=2)
        out["progressive_unet_3stage"] = {"ms_per_step": ms, "windows_per_s": B / ms * 1e3,
                                          "tflops": 3 * 288.627 * B / ms}
    if "deepcnn" in which:
        dm = b200sr.DeepCNN()
        dtr = b200sr.DeepCNNTrainer(dm, device=dev, model_save_dir="/tmp/b200sr_v", verbose=False)
        gen = b200sr.SyntheticTripletGenerator(B, 256, 256, device=dev, seed=1)
        x, y = gen.next()
        ms = timed(lambda: dtr.train_step(x, y), steps=4, warmup=2)
        # 1463 GFLOP forward per sample, x3 for the train step
        out["deepcnn_train"] = {"ms_per_step": ms, "triplets_per_s": B / ms * 1e3, "tflops": 3 * 1463.0 * B / ms}
        dm.eval()
        with torch.no_grad():
            ms = timed(lambda: dm(x), steps=4, warmup=2)
        out["deepcnn_infer"] = {"ms_per_batch": ms, "triplets_per_s": B / ms * 1e3, "tflops": 1463.0 * B / ms}
    if "fastddpm" in which:
        # BASELINE configs[4]: denoiser train step and T=10 DDIM sampling at 256x256. Algorithmic FLOPs are the
        # reference's (77.83 GFLOP fwd/sample incl. the 259-channel first conv); the engine folds the tiled time channels
        # into a bias table and issues 58.5 of them.
        fm = b200sr.FastDDPM(T=10, device=dev)
        ftr = b200sr.FastDDPMTrainer(fm, device=dev, model_save_dir="/tmp/b200sr_v", verbose=False)
        gen = b200sr.SyntheticTripletGenerator(B, 256, 256, device=dev, seed=1)
        x, y = gen.next()
        ms = timed(lambda: ftr.train_step(x, y), steps=10, warmup=3)
        out["fastddpm_train"] = {"ms_per_step": ms, "triplets_per_s": B / ms * 1e3, "tflops_algorithmic": 3 * 77.83 * B / ms,
                                 "tflops_issued": 3 * 58.5 * B / ms}
        fm.eval()
        ms = timed(lambda: fm.sample(x, dev), steps=6, warmup=4)  # the third call per shape captures the CUDA graph
        out["fastddpm_sample_T10"] = {"ms_per_batch": ms, "slices_per_s": B / ms * 1e3,
                                      "denoiser_evals_per_s": 10 * B / ms * 1e3, "tflops_algorithmic": 10 * 77.83 * B / ms}
        if os.environ.get("FD_PROFILE"):
            from b200sr import _lib
            _lib.enable_profiling(True)
            fm.train()
            for _ in range(3):
                ftr.train_step(x, y)
            agg = _lib.collect_profile()
            _lib.enable_profiling(False)
            out["fastddpm_train_per_op"] = {k: {"ms_per_step": v["ms"] / 3, "n": v["n"] // 3,
                                                "tflops": v["flop"] / (v["ms"] / 1e3) / 1e12 if v["flop"] else None}
                                            for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    return out


if __name__ == "__main__":
    print(json.dumps(run(tuple(sys.argv[1:]) or ("perceptual", "progressive", "deepcnn", "fastddpm"),
                         int(os.environ.get("B", "32")))))
