"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pair sharding, the max-over-ranks timing
reduction of the bench contract, the bucketed gradient all-reduce of the training step and the rank-0 broadcast of module buffers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from airpose_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_range(2048, 8, 3) == (768, 1024)          # BASELINE config 5: 256 pairs per GPU
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) sharding: each rank takes its slice of a global batch of pairs; together they cover it
        g = torch.Generator().manual_seed(0)
        batch = {"im0": torch.randn(7, 3, 4, 4, generator=g), "bb0": torch.randn(7, 3, generator=g)}
        mine = parallel.shard_batch(batch, world, rank)
        b, e = parallel.shard_range(7, world, rank)
        cover = torch.zeros(7, 3)
        cover[b:e] = mine["bb0"]                     # uneven shares (4 + 3): place, then sum over ranks
        dist.all_reduce(cover)
        ok_shard = torch.equal(cover, batch["bb0"]) and mine["im0"].shape[0] == e - b
        # (2) bench timing: max over ranks
        ms = parallel.max_over_ranks([1.0 + rank, 5.0 - rank])
        ok_max = ms == [float(world), 5.0]
        # (3) gradient all-reduce: three parameters, one without a gradient on rank 1, tiny buckets
        torch.manual_seed(1)
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
        params[0].grad = torch.full((5, 3), float(rank + 1))
        params[1].grad = torch.arange(7.0) * (rank + 1)
        if rank == 0:
            params[2].grad = torch.ones(2, 2) * 4
        ncoll = parallel.allreduce_mean_(params, bucket_bytes=64)
        mean_scale = sum(r + 1 for r in range(world)) / world
        ok_grad = (torch.allclose(params[0].grad, torch.full((5, 3), mean_scale)) and
                   torch.allclose(params[1].grad, torch.arange(7.0) * mean_scale) and
                   torch.allclose(params[2].grad, torch.ones(2, 2) * 4 / world))
        # (4) initial state: every rank starts from rank 0's buffers (what DDP's constructor does); float and integer buffers
        torch.manual_seed(10 + rank)
        bn = torch.nn.BatchNorm1d(6)
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.num_batches_tracked.fill_(3 + rank)
        parallel.broadcast_buffers_(bn)
        chk = torch.cat([bn.running_mean, bn.running_var, bn.num_batches_tracked.float().view(1)])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok_buf = torch.equal(lo, hi) and int(bn.num_batches_tracked) == 3
        # (5) the two-part gradient all-reduce of copenet_twoview.training_step: the part behind the split is reduced first
        # (asynchronously, while the "late" gradients are still being written), the front part afterwards; together they must
        # equal one all-reduce of the whole flat buffer
        offsets, numels = [0, 8, 20, 24], [6, 10, 3, 9]                 # 4-element slots, as optim.Adam lays them out
        split = parallel.late_split_offset(offsets, numels, [True, True, False, False])
        ok_two = split == 20 and parallel.late_split_offset(offsets, numels, [False] * 4) == 0
        flat = torch.arange(36.0) * (rank + 1)
        whole = flat.clone()
        dist.all_reduce(whole)
        flat[:split] = -1.0                                              # not final yet when the first part starts
        w1 = parallel.allreduce_begin(flat, split)
        flat[:split] = torch.arange(float(split)) * (rank + 1)           # the late gradients arrive
        w2 = parallel.allreduce_begin(flat, 0, split)
        for w in (w1, w2):
            w.wait()
        ok_two = ok_two and torch.equal(flat, whole) and parallel.allreduce_begin(flat, 5, 5) is None
        q.put((rank, ok_shard, ok_max, ok_grad and ok_buf and ok_two, ncoll))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_shard, ok_max, ok_grad, ncoll in res:
        assert ok_shard, "rank %d: shards do not tile the batch" % rank
        assert ok_max, "rank %d: max-over-ranks reduction wrong" % rank
        assert ok_grad, "rank %d: gradient mean wrong" % rank
        assert ncoll == 2                      # 60 B | 28 B + 16 B gradients with 64-byte buckets -> 2 collectives
