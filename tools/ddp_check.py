"""Multi-GPU sanity check (torchrun, one process per GPU, NCCL): after a few data-parallel steps on DIFFERENT per-rank
batches every rank must hold bit-identical weights (same all-reduced gradients, same Adam update), for the UNet
trainer and the Fast-DDPM trainer; and the sharded Fast-DDPM sampler (no collective) must reproduce the single-process
result for its shard.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import b200sr
from b200sr.ddp import shard_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def spread(model):
    """max over parameters of (max over ranks - min over ranks)"""
    worst = 0.0
    for p in model.parameters():
        hi, lo = p.detach().clone(), p.detach().clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        worst = max(worst, float((hi - lo).abs().max()))
    return worst


torch.manual_seed(100 + rank)  # different initial weights per rank: the trainer must broadcast rank 0's
unet = b200sr.UNet()
tr = b200sr.UNetTrainer(unet, device=dev, loss="combined", model_save_dir="/tmp/b200sr_ddp", verbose=False)
gen = b200sr.SyntheticTripletGenerator(4, 128, 256, device=dev, seed=7, rank=rank)
losses = [float(tr.train_step(*gen.next())) for _ in range(3)]
s_unet = spread(unet)

torch.manual_seed(200 + rank)
ddpm = b200sr.FastDDPM(T=10, device=dev)
ftr = b200sr.FastDDPMTrainer(ddpm, device=dev, model_save_dir="/tmp/b200sr_ddp", verbose=False)
gen2 = b200sr.SyntheticTripletGenerator(4, 64, 64, device=dev, seed=9, rank=rank)
flosses = [float(ftr.train_step(*gen2.next())) for _ in range(3)]
s_ddpm = spread(ddpm)

# sampling shards the batch with no collective: every rank samples its slice of the same global (cond, x_T)
g = torch.Generator(device="cpu").manual_seed(5)
cond = torch.randn(4 * world, 2, 64, 64, generator=g).to(dev)
x_T = torch.randn(4 * world, 1, 64, 64, generator=g).to(dev)
lo, hi = shard_batch(cond.shape[0], rank, world)
ddpm.eval()
mine = ddpm.sample(cond[lo:hi], dev, noise=x_T[lo:hi])
full = ddpm.sample(cond, dev, noise=x_T)[lo:hi]
s_sample = float((mine - full).abs().max())

ok = s_unet == 0.0 and s_ddpm == 0.0 and s_sample < 1e-5 and all(l == l for l in losses + flosses)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: weight spread after 3 steps unet {s_unet:.3e}, fastddpm {s_ddpm:.3e}; sharded-vs-full sample "
          f"{s_sample:.3e}; losses {losses} / {flosses}; {'DDP CHECK OK' if flag.item() == 1.0 else 'DDP CHECK FAILED'}")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
