"""Multi-rank numerical parity of the data-parallel train step: world-size-2 run, each rank its own
half of a global batch, BatchNorm statistics per rank (plain DDP semantics, no SyncBN), gradients all-reduced by the
bucketed reducer overlapped with backward — compared with the CPU ORACLE RUN PER SHARD AND AVERAGED.

Two ranks are spawned from the test. With two or more GPUs visible they use NCCL on separate devices (the production
path). With a single GPU (the driver's GPU test box) both ranks share cuda:0 and the collective runs on gloo — NCCL
refuses two ranks on one device — which still exercises everything that is this repo's: the engine, the per-block bucket
ranges, the event ordering between the backward streams and the communication stream, 1/world folded into Adam.
"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = dict(B=2, H=256, W=256, seed=4242)   # per-rank batch 2: global batch 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ngpu, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dev = torch.device("cuda", rank if ngpu >= world else 0)
    torch.cuda.set_device(dev)
    if ngpu >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200sr
    from oracle import cases
    torch.manual_seed(100 + rank)   # different initial weights per rank: the trainer must broadcast rank 0's
    model = b200sr.UNet()
    if rank == 0:
        model.load_state_dict(cases.seeded_state_dict(b200sr.UNet))
    tr = b200sr.UNetTrainer(model, device=dev, loss="combined", ssim_weight=0.005, learning_rate=1e-4,
                            model_save_dir="/tmp/b200sr_ddp_test", verbose=False)
    x, y = cases.seeded_batch(world * CASE["B"], CASE["H"], CASE["W"], CASE["seed"])
    lo, hi = rank * CASE["B"], (rank + 1) * CASE["B"]
    eng = model._get_engine()
    # the step, phase by phase, exactly as UNetTrainer._device_step issues it (so the reduced gradients can be read)
    tr.model.train()
    tr.optimizer.host_pre_step()
    out = eng.forward_train(x[lo:hi].to(dev))
    loss, dout = tr.criterion.value_and_grad(out, y[lo:hi].to(dev))
    from b200sr.ddp import BucketReducer
    red = BucketReducer(eng.flat_g)
    eng.backward(dout, bucket_hook=red.reduce_range)
    launched = list(red.launched)
    red.wait()
    torch.cuda.synchronize()
    summed = eng.flat_g.clone()
    tr.optimizer.device_step(grad_scale=1.0 / world)
    torch.cuda.synchronize()
    names = [n for n, _ in model.named_parameters()]
    grads = {n: g.detach().cpu().clone() / world for n, g in zip(names, eng.grad_views)}
    weights = {n: p.detach().cpu().clone() for n, p in model.named_parameters()}
    # a second full step through the public API (UNetTrainer.train_step) keeps the ranks in lock-step
    l2 = float(tr.train_step(x[lo:hi].to(dev), y[lo:hi].to(dev)))
    w2 = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu()
    # numpy payloads: torch tensors on a multiprocessing queue are passed by file descriptor and need the sender alive
    q.put((rank, float(loss), {k: v.numpy() for k, v in grads.items()} if rank == 0 else None,
           {k: v.numpy() for k, v in weights.items()} if rank == 0 else None, launched, w2.numpy(), l2,
           float(summed.abs().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_match_oracle_per_shard_average():
    world = 2
    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ngpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import b200sr
    from oracle import cases, ssim_oracle, unet_oracle
    sd = cases.seeded_state_dict(b200sr.UNet)
    x, y = cases.seeded_batch(world * CASE["B"], CASE["H"], CASE["W"], CASE["seed"])
    loss_fn = lambda p, t: ssim_oracle.combined_loss(p, t, 1.0, 0.005, "gaussian")
    o_losses, o_grads = [], None
    for r in range(world):   # the oracle per shard (its own BatchNorm statistics), gradients averaged
        sl = slice(r * CASE["B"], (r + 1) * CASE["B"])
        l, _, g, _ = unet_oracle.loss_and_grads(sd, x[sl], y[sl], loss_fn)
        o_losses.append(float(l))
        o_grads = g if o_grads is None else {k: o_grads[k] + g[k] for k in g}
    o_grads = {k: v / world for k, v in o_grads.items()}

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))

    for r in range(world):
        assert abs(results[r][1] - o_losses[r]) / abs(o_losses[r]) < 1e-3, (r, results[r][1], o_losses[r])
    grads = {k: torch.from_numpy(v) for k, v in results[0][2].items()}
    weights = {k: torch.from_numpy(v) for k, v in results[0][3].items()}
    worst = {}
    for n, g in grads.items():
        if n.endswith("conv.0.bias") or n.endswith("conv.3.bias"):
            assert float(g.abs().max()) == 0.0
            continue
        r_, c_ = rel(g, o_grads[n]), cos(g, o_grads[n])
        worst[n] = (r_, c_)
        # same structure of gates as the single-rank tests (per-tensor calibration exists for B=2 128x256 and B=32
        # 256x256; this B=2+2 case sits between them, so the bounds are the loosest calibrated ones)
        shallow = n.startswith(("final_conv", "dec1"))
        assert c_ >= (0.99 if shallow else 0.75), (n, r_, c_)
        assert r_ <= (5e-2 if shallow else 0.75), (n, r_, c_)
    # Adam consumed the all-reduced SUM with grad_scale 1/world: weights == torch Adam on the averaged gradient
    for n, w in weights.items():
        exp, _, _ = unet_oracle.adam_update(sd[n].double(), grads[n].double(), 0.0, 0.0, 1)
        assert float((w.double() - exp).abs().max()) <= 2e-7, n
    # the reducer saw one range per block group and launched several buckets before the final flush
    assert len(results[0][4]) >= 3, results[0][4]
    # both ranks hold bit-identical weights after two steps
    assert (results[0][5] == results[1][5]).all()
    assert results[0][7] > 0
