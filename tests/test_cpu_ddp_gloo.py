"""world_size-2 gloo test of the data-parallel plumbing (BucketReducer, broadcast) on CPU tensors."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import b200sr
    from b200sr.ddp import BucketReducer, broadcast_module_state, is_distributed
    assert is_distributed()
    # 1. parameters and buffers start identical on all ranks
    torch.manual_seed(100 + rank)
    m = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3), torch.nn.BatchNorm2d(4))
    m[1].running_mean.fill_(float(rank))
    broadcast_module_state(m)
    ref = [t.clone() for t in list(m.parameters()) + list(m.buffers())]
    gathered = [None] * world
    dist.all_gather_object(gathered, [t.tolist() for t in ref])
    same = all(g == gathered[0] for g in gathered)
    # 2. bucketed reduction of a flat gradient buffer reported in descending ranges
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red = BucketReducer(flat, min_bucket_elems=300)
    for lo, hi in ((900, 1000), (600, 900), (590, 600), (100, 590), (0, 100)):
        red.reduce_range(lo, hi)
    launched_before_wait = list(red.launched)
    red.wait()
    expect = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
    q.put((rank, same, bool(torch.equal(flat, expect)), launched_before_wait))
    dist.destroy_process_group()


def test_bucket_reducer_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, ok, launched in results:
        assert same, "broadcast_module_state left ranks different"
        assert ok, "bucketed all-reduce produced a wrong sum"
        assert launched == [(600, 1000), (100, 600)], launched  # (0,100) is flushed by wait()
