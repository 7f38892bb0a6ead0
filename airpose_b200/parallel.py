"""Multi-GPU plumbing of the hot path: one process per GPU, ``torch.distributed`` for the control plane.

The copenet_twoview forward shards by frame pair with NO data-path collective: every
pair is independent in eval mode and the two views of a pair stay on one rank (they exchange 136 floats
per regressor iteration, model_copenet.py:185,192).  What the reference gets from Lightning's DDP wrapper
(copenet_trainer.py:56-65) and what a drop-in needs from this module is therefore small:

* ``shard_range``      contiguous split of a global batch of pairs over the ranks (config 5: 2048 -> 256/GPU)
* ``max_over_ranks``   device-side timing reduction of the bench contract (max over ranks, never wall clock)
* ``allreduce_mean_``  the one collective of the training step (config 4): gradient mean over ranks, in
                       flat buckets so that launch latency is paid per bucket rather than per parameter;
                       parameters without a gradient (``deccam`` in the two-view model, model_copenet.py:73)
                       contribute zeros, which is what DDP's find_unused_parameters does.

Backend: NCCL over NVLink on the GPU box, gloo in the CPU tests (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous share of ``n_items``; the first ``n_items % world_size`` ranks get one
    more item, every item belongs to exactly one rank, empty shares are legal (n_items < world_size)."""
    if world_size < 1 or not 0 <= rank < world_size or n_items < 0:
        raise ValueError("bad shard request n=%d world=%d rank=%d" % (n_items, world_size, rank))
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(batch: dict, world_size: int, rank: int) -> dict:
    """Slice every tensor of a batch dict (same leading dimension = pairs) to this rank's share."""
    n = next(iter(batch.values())).shape[0]
    b, e = shard_range(n, world_size, rank)
    return {k: v[b:e] for k, v in batch.items()}


def max_over_ranks(values: Sequence[float], device=None, group=None) -> List[float]:
    """Elementwise max of a few per-rank scalars (step times in ms) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.tolist()


def _buckets(tensors: Sequence[torch.Tensor], bucket_bytes: int) -> Iterable[List[torch.Tensor]]:
    cur, size = [], 0
    for t in tensors:
        nbytes = t.numel() * t.element_size()
        if cur and (size + nbytes > bucket_bytes or t.dtype != cur[0].dtype):
            yield cur
            cur, size = [], 0
        cur.append(t)
        size += nbytes
    if cur:
        yield cur


def allreduce_mean_(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None) -> int:
    """In-place mean of ``p.grad`` over all ranks for every parameter, bucketed.  Parameters whose grad is
    None on this rank take part with zeros (and receive the mean of the others), so ranks never disagree
    on the bucket layout.  Returns the number of collectives issued.  64 MiB buckets: 27.1 M fp32
    parameters (108 MB) go out in two all-reduces; over NVSwitch the cost is launch latency, not links."""
    params = [p for p in params if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    n = 0
    for bucket in _buckets([p.grad for p in params], bucket_bytes):
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
    return n


def broadcast_buffers_(module, src=0, group=None):
    """Rank ``src``'s module buffers (BatchNorm running statistics, ``num_batches_tracked``, mean-pose buffers) to every rank,
    as DistributedDataParallel does at construction (``broadcast_buffers``).  One flat broadcast per dtype."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return module
    by_dtype = {}
    for b in module.buffers():
        by_dtype.setdefault(b.dtype, []).append(b)
    for bufs in by_dtype.values():
        flat = torch.cat([b.detach().reshape(-1) for b in bufs])
        dist.broadcast(flat, src=src, group=group)
        o = 0
        with torch.no_grad():
            for b in bufs:
                b.copy_(flat[o:o + b.numel()].view_as(b))
                o += b.numel()
    return module


def late_split_offset(offsets, numels, late):
    """Where the flat gradient buffer of ``optim.Adam`` is cut for the overlapped all-reduce: the end (rounded up to the 4-element
    slot pitch) of the last parameter flagged ``late`` -- a parameter whose gradient is only final at the end of the backward
    (conv1 / bn1 / layer1 / layer2).  Everything from that offset on is reduced while the rest of the backward still runs."""
    end = 0
    for o, n, is_late in zip(offsets, numels, late):
        if is_late:
            end = max(end, o + (n + 3) // 4 * 4)
    return end


def allreduce_begin(flat, lo, hi=None, group=None):
    """Asynchronous SUM all-reduce of ``flat[lo:hi]``, ordered after the work already enqueued on the current stream (NCCL) /
    issued immediately (gloo); returns the work handle, or None when there is nothing to reduce (one rank, empty range)."""
    hi = flat.numel() if hi is None else hi
    if hi <= lo or not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return None
    return dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)

