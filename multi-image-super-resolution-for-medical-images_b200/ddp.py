"""Data-parallel plumbing for the UNet train step: one process per GPU, torch.distributed (NCCL over NVLink).

The reference has no multi-GPU path (only nn.DataParallel in one FastDDPM notebook); this is the batch-sharded
DDP the north star asks for. The engine's backward writes every gradient into ONE flat fp32 buffer and reports
ranges of it as soon as they are final (reverse forward order). `BucketReducer` launches one asynchronous
all-reduce(sum) per reported range on a side stream, ordered behind the producing kernels by a CUDA event, so the
wire time overlaps the rest of backward. The 1/world_size factor is folded into the optimizer's grad_scale.

BatchNorm statistics stay per-rank (plain DDP semantics, no SyncBN); buffers are broadcast from rank 0 once.
On CPU tensors (gloo; used by the world_size-2 tests) the same code runs synchronously.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_module_state(module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not is_distributed():
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)


def alloc_comm_buffer(numel: int, device):
    """A zeroed flat fp32 buffer allocated by NCCL's own allocator (ncclMemAlloc) and registered with the communicator
    (user-buffer registration): all-reduces on (ranges of) it can then run zero-copy, with in-switch reduction (NVLS) on
    NVSwitch systems, instead of staging through NCCL's internal buffers — fewer SMs and less HBM traffic taken from the
    backward pass they overlap. Returns None when not distributed / not NCCL / not supported by this torch build
    (B200SR_NO_NCCL_REG=1 switches it off); the caller then falls back to an ordinary allocation."""
    import os
    if not is_distributed() or os.environ.get("B200SR_NO_NCCL_REG") is not None:
        return None
    try:
        if dist.get_backend() != "nccl":
            return None
        backend = dist.group.WORLD._get_backend(torch.device(device))
        pool = torch.cuda.MemPool(backend.mem_allocator)
        with torch.cuda.use_mem_pool(pool):
            buf = torch.zeros(numel, dtype=torch.float32, device=device)
        backend.register_mem_pool(pool)
        buf._b200sr_pool = pool  # keep the pool (and its registration) alive as long as the buffer
        return buf
    except Exception:  # older torch / NCCL, or a backend without a memory allocator
        return None


def shard_batch(n_total: int, rank: int, world_size: int):
    """Contiguous [lo, hi) slice of a global batch owned by `rank` (inference shards with no collective)."""
    per = (n_total + world_size - 1) // world_size
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


class BucketReducer:
    """All-reduces ranges of a flat gradient buffer as they become final.

    min_bucket_elems coalesces small trailing ranges: a reported range is deferred until at least that many
    elements are pending (the last call of a step, `flush()`, sends whatever is left)."""

    def __init__(self, flat: torch.Tensor, group=None, min_bucket_elems: int = 1 << 20):
        self.flat = flat
        self.group = group
        self.min_bucket_elems = min_bucket_elems
        self.world_size = dist.get_world_size(group) if is_distributed() else 1
        self._pending = None  # (lo, hi) not yet sent
        self._works = []
        self.launched = []  # (lo, hi) of every collective of the current step, for tests / accounting
        self._cuda = flat.is_cuda
        if self._cuda:
            self.comm_stream = torch.cuda.Stream(device=flat.device)
            self._event = torch.cuda.Event()

    def _launch(self, lo: int, hi: int) -> None:
        if hi <= lo or self.world_size == 1:
            return
        self.launched.append((lo, hi))
        view = self.flat[lo:hi]
        if self._cuda:
            self._event.record(torch.cuda.current_stream(self.flat.device))
            self.comm_stream.wait_event(self._event)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_range(self, lo: int, hi: int) -> None:
        """flat[lo:hi] is final on the current stream. Ranges arrive in descending, adjacent order."""
        if self._pending is not None and self._pending[0] == hi:
            lo, hi = lo, self._pending[1]
        elif self._pending is not None:
            self._launch(*self._pending)
        self._pending = (lo, hi)
        if hi - lo >= self.min_bucket_elems:
            self._launch(lo, hi)
            self._pending = None

    def flush(self) -> None:
        if self._pending is not None:
            self._launch(*self._pending)
            self._pending = None

    def wait(self) -> None:
        """Make the current stream wait for every collective launched this step."""
        self.flush()
        if self._cuda and self.world_size > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
        self.launched = []
