#!/usr/bin/env python
"""Benchmark of the b200sr hot path: UNet 2->1 slice triplets/sec @256^2 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload at every N: BASELINE.json configs[2] restricted to the hot path the north star names — UNet train step
with the combined MSE + 0.005*(1-SSIM) loss and Adam(lr 1e-4), bf16 tensor-core compute, batch 32 per GPU at
256x256, data-parallel by batch (weak scaling; gradient all-reduce over NCCL overlapped with backward). Inference
throughput (configs[0]: B=8 fp32 in/out, eval mode) is reported in the same JSON line under "inference"; at N=1 the
other BASELINE configs built on the same kernels (the step with the VGG16 perceptual term, the Progressive
UNet chain, the DeepCNN baseline, Fast-DDPM training and 10-step sampling) are timed under "variants".

Prints ONE JSON line (rank 0). See DESIGN.md §Measurement for how every field is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

# stdout carries exactly ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1 behind Python's back, so
# fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved original descriptor
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_TRIPLET = 288.627   # fwd + dgrad + wgrad, 2*MACs
FWD_GFLOP_PER_TRIPLET = 96.259
METRIC = "unet_train_triplets_per_sec_256x256"
UNIT = "triplets/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic_per_launch():
    """Mean DRAM bytes (read + write) per tensor-core launch of one B=32 train step, parsed from the committed ncu
    summary of the same launch population (tools/ncu_step.py under ncu --set full; tools/ncu_summary.py)."""
    path = os.path.join(ROOT, "profiles", "r2_gemm_ncu_summary.txt")
    if not os.path.exists(path):
        return None, "profiles/r2_gemm_ncu_summary.txt missing"
    tot, n = 0.0, 0
    with open(path) as f:
        for line in f:
            if line.startswith("#") or "dram_rd_MB=" not in line:
                continue
            try:
                rd = float(line.split("dram_rd_MB=")[1].split()[0])
                wr = float(line.split("dram_wr_MB=")[1].split()[0])
            except (IndexError, ValueError):
                continue
            tot += (rd + wr) * 1e6
            n += 1
    if n == 0:
        return None, "no launches parsed from profiles/r2_gemm_ncu_summary.txt"
    return tot / n, f"profiles/r2_gemm_ncu_summary.txt: {n} tensor-core launches of one train step (ncu --set full, cold cache)"


def bench_config(world, B):
    """The workload both arms are quoted on (the reference arm times a bounded sample of it, see cpu_baseline.sample)."""
    return {"workload": "unet_train_combined_mse_ssim_b32_256x256 (BASELINE configs[2] hot path: UNet fwd+bwd, "
                        "MSE+0.005*(1-SSIM), Adam; the same step with the VGG16 perceptual term is reported as "
                        "'full_combined_loss', the other BASELINE configs under 'variants')",
            "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
            "l2": "per-step working set (activations + gradients, several GB) >> 126 MB L2; ring of 4 distinct input batches"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port; the reference is Python/PyTorch and cannot travel to the
# GPU box) on the host CPU cores
# ----------------------------------------------------------------------------------------------------------
def cpu_train_sample(batch, steps, warmup, threads=None):
    import torch
    import b200sr
    from oracle import cases, ssim_oracle, unet_oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = cases.seeded_state_dict(b200sr.UNet)
    x, y = cases.seeded_batch(batch, 256, 256, 1234)
    loss_fn = lambda p, t: ssim_oracle.combined_loss(p, t, 1.0, 0.005, "gaussian")
    opt_state = {}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        unet_oracle.train_step_cpu(sd, x, y, opt_state, i + 1, loss_fn=loss_fn)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * steps / total, "ms_per_step": 1e3 * total / steps, "cores": threads,
            "sample": f"bounded sample of the workload: {steps} train steps (UNet fwd+bwd, MSE+0.005*(1-SSIM), Adam) at "
                      f"batch {batch} instead of 32 (throughput in triplets/s is batch-size normalised), 256x256 fp32, "
                      f"torch {torch.__version__} CPU, {warmup} warm-up"}


def cpu_infer_sample(batch=8, reps=2, threads=None):
    """BASELINE configs[0] on the host: UNet eval forward (B=8,2,256,256) fp32 through the oracle port, best of `reps`
    after one warm-up."""
    import torch
    import b200sr
    from oracle import cases, unet_oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = cases.seeded_state_dict(b200sr.UNet)
    x, _ = cases.seeded_batch(batch, 256, 256, 4321)
    best = None
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            unet_oracle.unet_forward(sd, x, training=False)
            dt = time.perf_counter() - t0
            if i > 0:
                best = dt if best is None else min(best, dt)
    return {"value": batch / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"best of {reps} eval forwards at batch {batch}, 256x256 fp32, torch {torch.__version__} CPU, 1 warm-up"}




# ----------------------------------------------------------------------------------------------------------
# library baseline on the SAME GPU: the reference algorithm as stock PyTorch ops (ATen / cuDNN), i.e. what a user of the
# reference gets on this box today. It is the oracle's functional restatement of the reference UNet (pinned bit-for-bit to
# the reference modules, oracle/make_golden.py) moved to `cuda`, channels_last, cudnn.benchmark on — a BASELINE leg like
# the CPU one: nothing of this repo's engine or kernels runs in it, and it is never the thing shipped.
# ----------------------------------------------------------------------------------------------------------
def gpu_library_baseline(dev, B, steps=8, warmup=3):
    import torch
    import b200sr
    from oracle import cases, ssim_oracle, unet_oracle
    cl = torch.channels_last
    sd0 = cases.seeded_state_dict(b200sr.UNet)
    x, y = cases.seeded_batch(B, 256, 256, 1234)
    x, y = x.to(dev).contiguous(memory_format=cl), y.to(dev)
    xe = x[:8].contiguous(memory_format=cl)
    names = set(unet_oracle.param_names(sd0))
    loss_fn = lambda p, t: ssim_oracle.combined_loss(p.float(), t, 1.0, 0.005, "gaussian")
    was = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    out = {}

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        for mode in ("bf16_autocast", "tf32", "fp32"):
            tf32 = mode != "fp32"
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            sd = {}
            for k, v in sd0.items():
                t = v.to(dev)
                if t.dim() == 4:
                    t = t.contiguous(memory_format=cl)
                sd[k] = t.requires_grad_(True) if k in names else t
            params = [sd[k] for k in unet_oracle.param_names(sd0)]
            opt = torch.optim.Adam(params, lr=1e-4, fused=True)
            ac = torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast"))

            def train():
                stats = {}
                with ac:
                    pred = unet_oracle.unet_forward(sd, x, training=True, new_stats=stats)
                loss = loss_fn(pred, y)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                sd.update(stats)   # running statistics of the step

            def infer():
                with torch.no_grad(), ac:
                    unet_oracle.unet_forward(sd, xe, training=False)

            ent = {}
            if mode != "fp32":
                ms = timed(train, steps, warmup)
                ent["train_ms_per_step"], ent["train_triplets_per_s"] = ms, B / (ms / 1e3)
            ms = timed(infer, 20, 5)
            ent["infer_b8_ms_per_batch"], ent["infer_b8_triplets_per_s"] = ms, 8 / (ms / 1e3)
            out[mode] = ent
            del opt, params, sd
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = was
    out["what"] = (f"stock PyTorch {torch.__version__} (ATen/cuDNN {torch.backends.cudnn.version()}) on the same GPU: reference "
                   f"UNet as plain torch ops, channels_last, cudnn.benchmark, fused Adam; train = fwd + MSE+0.005*(1-SSIM) + "
                   f"bwd + Adam at batch {B}, infer = eval forward at batch 8; modes: bf16 autocast, TF32, strict fp32")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 4
    r = cpu_train_sample(batch, args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(max(int(os.environ.get("WORLD_SIZE", "1")), 1), args.batch),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------
def run_b200sr(args):
    import torch
    import torch.distributed as dist
    import b200sr
    from b200sr import _lib
    from oracle import cases

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W = args.batch, 256, 256
    peaks = load_peaks()

    model = b200sr.UNet()
    model.load_state_dict(cases.seeded_state_dict(b200sr.UNet))
    trainer = b200sr.UNetTrainer(model, device=dev, loss="combined", ssim_weight=0.005, learning_rate=1e-4,
                                 model_save_dir="/tmp/b200sr_bench", verbose=False)
    gen = b200sr.SyntheticTripletGenerator(B, H, W, device=dev, seed=1234, rank=rank)
    ring = [gen.next() for _ in range(4)]
    host_ring = [(x.cpu().pin_memory(), y.cpu().pin_memory()) for x, y in ring]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: K train steps, inputs resident in HBM ------------------------------------------------------
    def step_resident(i):
        x, y = ring[i % len(ring)]
        trainer.train_step(x, y)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.LAUNCH_COUNTER["n"] = 0
    total_ms = timed(step_resident, args.steps, args.warmup)
    launches_total = _lib.LAUNCH_COUNTER["n"]
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)
    grad_buffer_registered = bool(getattr(model._get_engine(), "flat_g_registered", False))
    launches_timed = launches_total * args.steps // (args.steps + args.warmup)

    # ---- e2e: the public API from pinned HOST buffers: DevicePrefetcher (H2D of batch i+1 overlaps step i) feeding
    # UNetTrainer.train_step, loss read back to the host every step (like the reference's loss.item(), :187) --------
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def run_e2e(nsteps):
        # every step: H2D of its inputs (pinned host -> device, prefetched one step ahead) and a D2H read of its loss;
        # the host consumes the loss of step i-1 while step i runs (no per-step pipeline bubble)
        host_batches = (host_ring[i % len(host_ring)] for i in range(nsteps))
        for i, (x, y) in enumerate(b200sr.DevicePrefetcher(host_batches, dev)):
            loss = trainer.train_step(x, y)
            loss_host[i % 2:i % 2 + 1].copy_(loss.reshape(1), non_blocking=True)
            loss_ev[i % 2].record()
            if i > 0:
                loss_ev[1 - i % 2].synchronize()
                _ = float(loss_host[1 - i % 2])
        loss_ev[(nsteps - 1) % 2].synchronize()

    run_e2e(3)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    sync_all()
    e2e_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t.item())
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host_ring[0])

    # ---- inference (configs[0] shape: B=8 fp32 in/out, eval mode; batch-sharded, no collective) --------------
    model.eval()
    xi = ring[0][0][:8].contiguous()
    with torch.no_grad():
        inf_ms = timed(lambda i: model(xi), 20, 5)
        inf_value = world * 8 * 20 / (inf_ms / 1e3)
        xb = ring[0][0]
        inf_big_ms = timed(lambda i: model(xb), 10, 3)
        inf_big_value = world * B * 10 / (inf_big_ms / 1e3)
        # configs[0] at the reference's own precision (fp32 forward; bf16x3 operand splitting, rel-L2 <= 1e-4)
        model.set_eval_precision("fp32")
        inf32_ms = timed(lambda i: model(xi), 20, 5)
        inf32_value = world * 8 * 20 / (inf32_ms / 1e3)
        inf32_big_ms = timed(lambda i: model(xb), 10, 3)
        inf32_big_value = world * B * 10 / (inf32_big_ms / 1e3)
        model.set_eval_precision("bf16")

    # ---- fused MSE+SSIM loss kernel at a batch large enough to leave the launch-latency regime (SURVEY hard-part 7) ---
    ssim_big = None
    if rank == 0:
        try:
            pb, tb = torch.randn(512, 1, H, W, device=dev), torch.randn(512, 1, H, W, device=dev)
            crit = trainer.criterion
            for _ in range(3):
                crit.value_and_grad(pb, tb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                crit.value_and_grad(pb, tb)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            gbs = 512 * 786432 / (ms / 1e3) / 1e9
            ssim_big = {"batch": 512, "ms": ms, "gbs_algorithmic": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"],
                        "note": "786,432 B/triplet (pred + target read, gradient written); instruction-issue bound: >= 220 "
                                "FMA/pixel for 12 B/pixel caps an fp32 CUDA-core implementation at 0.30 of HBM (DESIGN.md §3)"}
            del pb, tb
        except Exception as exc:
            ssim_big = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    # ---- the same step with the full BASELINE configs[2] loss (MSE + 0.01*VGG16 perceptual + 0.005*(1-SSIM)) -------------
    full_loss = None
    try:
        model.train()
        tr_full = b200sr.UNetTrainer(model, device=dev, loss="combined_perceptual", ssim_weight=0.005, learning_rate=1e-4,
                                     model_save_dir="/tmp/b200sr_bench", verbose=False)

        def step_full(i):
            x, y = ring[i % len(ring)]
            tr_full.train_step(x, y)

        full_ms = timed(step_full, max(args.steps // 2, 5), 3) / max(args.steps // 2, 5)
        full_loss = {"value": world * B / (full_ms / 1e3), "unit": UNIT, "ms_per_step": full_ms,
                     "workload": "unet_train_mse_vgg16perceptual_ssim_b32_256x256 (BASELINE configs[2] with its full "
                                 "combined loss; VGG16 features[:16] with seeded random-init weights, parity unpinned: the "
                                 "reference notebook defining the term is missing from the snapshot)"}
        del tr_full
    except Exception as exc:  # never take the headline down
        full_loss = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- roofline of the dominant kernel (tensor-core implicit GEMM), timed live with CUDA events ------------
    # every rank runs the 3 instrumented steps (the train step contains collectives); rank 0 records them
    roofline = None
    model.train()
    engine = model._get_engine()
    overlap_was, graph_was = engine.overlap_wgrad, trainer.use_cuda_graph
    engine.overlap_wgrad = False   # per-kernel event timing needs the kernels serialised on one stream ...
    trainer.use_cuda_graph = False  # ... and launched eagerly
    if rank == 0:
        _lib.enable_profiling(True)
    for i in range(3):
        step_resident(i)
    sync_all()
    engine.overlap_wgrad, trainer.use_cuda_graph = overlap_was, graph_was
    if rank == 0:
        agg = _lib.collect_profile()
        _lib.enable_profiling(False)
        gemm = {k: v for k, v in agg.items() if k in _lib.GEMM_OPS}
        g_ms = sum(v["ms"] for v in gemm.values())
        g_flop = sum(v["flop"] for v in gemm.values())
        g_n = sum(v["n"] for v in gemm.values())
        all_ms = sum(v["ms"] for v in agg.values())
        achieved = g_flop / (g_ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        traffic, traffic_src = ncu_traffic_per_launch()
        roofline = {"bound": "tensor", "kernel": "conv3x3_kernel + wgrad3x3_kernel + wgrad_kernel (tcgen05 implicit GEMM: conv3x3 "
                    "fwd/dgrad/wgrad, ConvT fwd/dgrad/wgrad)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak,
                    # DRAM bytes per tensor-core launch: dram__bytes_read.sum + dram__bytes_write.sum of every tensor-core
                    # launch of ONE train step of this very workload (ncu --set full, committed summary), mean per launch
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + " sustained bf16 (kernels timed inside a long step); burst "
                                   f"{peaks['bf16_tflops']}",
                    "note": "per-kernel times from 3 instrumented eager steps (CUDA graph and wgrad side stream off)",
                    "launches_per_step": g_n // 3, "avg_launch_ms": g_ms / max(g_n, 1),
                    "gflop_per_launch": g_flop / max(g_n, 1) / 1e9, "share_of_step": g_ms / all_ms,
                    "per_op": {k: {"ms_per_step": v["ms"] / 3, "tflops": (v["flop"] / (v["ms"] / 1e3) / 1e12)
                                   if v["flop"] else None, "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9)
                                   if v["bytes"] else None, "launches": v["n"] // 3}
                               for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
                    "step_tflops": TRAIN_GFLOP_PER_TRIPLET * B / ms_per_step,
                    "step_frac_of_peak": TRAIN_GFLOP_PER_TRIPLET * B / ms_per_step / peak}

    # ---- the other BASELINE.json configs built on the same kernels, N=1 only: each is a train step
    # (or sampler run) at batch 32/GPU, 256x256, timed with CUDA events after warm-up. Reported next to the headline,
    # never mixed into it. ------------------------------------------------------------------------------------------
    variants = None
    if world == 1 and not args.no_variants:
        import gc
        import importlib.util
        del trainer, model, engine, ring, gen
        gc.collect()
        torch.cuda.empty_cache()
        spec = importlib.util.spec_from_file_location("b200sr_bench_variants", os.path.join(ROOT, "tools", "bench_variants.py"))
        bv = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bv)
        variants = {}
        for name in ("perceptual", "progressive", "deepcnn", "fastddpm"):
            try:
                variants.update(bv.run((name,), B))
            except Exception as exc:  # a variant must never take the headline line down with it
                variants[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            gc.collect()
            torch.cuda.empty_cache()

    # ---- N > 1: BASELINE configs[4] ("batch-sharded 8 GPUs"): Fast-DDPM denoiser training (gradient all-reduce over NCCL)
    # and T=10 sampling (batch sharded, no collective), every rank B samples; timed like the headline (barrier, max) -------
    if world > 1 and not args.no_variants:
        try:
            fm = b200sr.FastDDPM(T=10, device=dev)
            ftr = b200sr.FastDDPMTrainer(fm, device=dev, model_save_dir="/tmp/b200sr_bench", verbose=False)
            fgen = b200sr.SyntheticTripletGenerator(B, H, W, device=dev, seed=1, rank=rank)
            fx, fy = fgen.next()
            f_ms = timed(lambda i: ftr.train_step(fx, fy), 10, 3) / 10
            fm.eval()
            s_ms = timed(lambda i: fm.sample(fx, dev), 6, 4) / 6
            variants = {"fastddpm_train": {"ms_per_step": f_ms, "triplets_per_s": world * B / f_ms * 1e3,
                                           "parallelism": f"dp{world}", "per_gpu_batch": B},
                        "fastddpm_sample_T10": {"ms_per_batch": s_ms, "slices_per_s": world * B / s_ms * 1e3,
                                                "denoiser_evals_per_s": 10 * world * B / s_ms * 1e3,
                                                "parallelism": f"batch-sharded x{world}, no collective"}}
            del fm, ftr
        except Exception as exc:
            variants = {"fastddpm": {"error": f"{type(exc).__name__}: {exc}"[:300]}}

    # ---- stock PyTorch (cuDNN) on the same GPU, same run (rank 0, N=1 only) ---------------------------------------
    gpu_lib = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        torch.cuda.empty_cache()
        try:
            gpu_lib = gpu_library_baseline(dev, B)
            for mode, ent in gpu_lib.items():
                if isinstance(ent, dict) and "train_triplets_per_s" in ent:
                    ent["this_repo_over_library_train"] = value / ent["train_triplets_per_s"]
                if isinstance(ent, dict) and "infer_b8_triplets_per_s" in ent:
                    mine = inf32_value if mode == "fp32" else inf_value
                    ent["this_repo_over_library_infer_b8"] = mine / ent["infer_b8_triplets_per_s"]
        except Exception as exc:
            gpu_lib = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) -------------------------------------------------------
    cpu, cpu_inf = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_train_sample(4, 3, 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        cpu_inf = cpu_infer_sample()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": bench_config(world, B),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches_timed, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "ddp": {"world": world, "collective": "NCCL all-reduce(sum) of the flat fp32 gradient buffer, one range per "
                        "UNet block, overlapped with backward" if world > 1 else None,
                        "nccl_user_buffer_registration": grad_buffer_registered},
                "full_combined_loss": full_loss, "gpu_library_baseline": gpu_lib, "ssim_kernel_b512": ssim_big,
                "variants": variants,
                "inference": {"value": inf_value, "unit": UNIT, "batch_per_gpu": 8, "ms_per_batch": inf_ms / 20,
                              "workload": "BASELINE configs[0]: UNet eval forward (B=8,2,256,256)->(B,1,256,256), "
                                          "fp32 in/out, bf16 tensor-core compute",
                              "dtype": "bf16", "value_b32": inf_big_value, "frac_of_peak_b32":
                                  FWD_GFLOP_PER_TRIPLET * inf_big_value / world / 1e3 / peaks["bf16_tflops"],
                              "fp32_mode": {"value": inf32_value, "unit": UNIT, "batch_per_gpu": 8, "dtype": "f32",
                                            "ms_per_batch": inf32_ms / 20, "value_b32": inf32_big_value,
                                            "workload": "BASELINE configs[0] at the reference's precision: fp32 in/out, fp32 "
                                                        "activations and accumulation via bf16x3 operand splitting on the "
                                                        "tensor cores (rel-L2 <= 1e-4 vs the fp32 oracle, "
                                                        "tests/test_gpu_parity_big.py)",
                                            "issued_tflops_b32": 3 * FWD_GFLOP_PER_TRIPLET * inf32_big_value / world / 1e3},
                              "cpu_baseline": cpu_inf}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200sr", choices=["b200sr", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (BASELINE configs[2]: 32)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--no-variants", action="store_true", help="skip the other BASELINE configs (N=1 only)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-on-the-same-GPU comparator")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200sr" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200sr(args)


if __name__ == "__main__":
    main()
