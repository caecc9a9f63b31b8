"""world_size-2 gloo (CPU) tests of the data-parallel host logic of pytorch_glow_b200/train.py: the flat
parameter / gradient arenas, gradient averaging before clipping, rank-0 initialisation broadcast and batch
sharding (reference: DataParallel scatter / reduce-add, network/trainer.py:117-123,138-150).
No CUDA: gradients are synthetic and the optimizer step is the oracle's (checker only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import glow_oracle as O
        from pytorch_glow_b200.train import FlatArena, allreduce_mean_, shard_batch
        torch.manual_seed(1234)                      # identical replicas on every rank
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
        model.register_parameter("h_top", torch.nn.Parameter(torch.zeros(2, 3)))
        arena = FlatArena(model)
        assert "h_top" not in [n.split(".")[-1] for n in arena.names]
        assert all(p.data_ptr() == arena.flat.data_ptr() + 4 * o for p, o in zip(arena.params, arena.offsets))
        # rank 0's (data-dependent) initialisation wins: trainer.py:112-115
        if rank == 0:
            arena.flat.add_(1.0)
        dist.broadcast(arena.flat, src=0)
        # this rank's shard of the global batch, rank-dependent gradients
        xg = torch.arange(8 * 7, dtype=torch.float32).reshape(8, 7) / 50
        x = shard_batch(xg, rank, world)
        assert x.shape[0] == 8 // world and torch.equal(x, xg[rank * 4:(rank + 1) * 4])
        loss = model(x).pow(2).mean()
        full = model(xg).pow(2).mean().detach()       # what a single process would optimise (trainer.py:126)
        arena.grad.zero_(); arena.rebind_grads()
        loss.backward()
        assert all(p.grad.data_ptr() == arena.grad.data_ptr() + 4 * o for p, o in zip(arena.params, arena.offsets))
        local = arena.grad.clone()
        allreduce_mean_(arena.grad)
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(arena.grad, sum(gathered) / world, rtol=0, atol=1e-7)
        # clip + Adam on identical averaged gradients => identical parameters on every rank, no second collective
        O.clip_grads_([arena.grad], 5.0, 100.0)
        m, v = torch.zeros_like(arena.flat), torch.zeros_like(arena.flat)
        with torch.no_grad():
            O.adam_step_(arena.flat, arena.grad, m, v, 1, O.noam_lr(1e-3, 0, 4000, 1e-4))
        flats = [torch.zeros_like(arena.flat) for _ in range(world)]
        dist.all_gather(flats, arena.flat)
        assert all(torch.equal(flats[0], f) for f in flats)
        # the mean of per-rank means over equal shards is the global-batch mean the reference optimises
        l = loss.detach().clone()
        allreduce_mean_(l)
        out[rank] = (float(l), float(full))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        assert out[r][0] == pytest.approx(out[0][0], abs=1e-7)          # every rank sees the same global loss
        assert out[r][0] == pytest.approx(out[r][1], rel=1e-6)          # = the un-sharded batch mean


def test_shard_batch_rejects_ragged_batches():
    from pytorch_glow_b200.train import shard_batch
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(7, 3), 0, 2)


def _worker_levels(rank, world, port, out):
    """Per-level (bucketed, asynchronous) gradient averaging == one all-reduce of the whole arena."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        import pytorch_glow_b200 as G
        from pytorch_glow_b200.hps import make_hps
        from pytorch_glow_b200.train import FlatArena, allreduce_mean_, arena_level_ranges
        np.random.seed(0); torch.manual_seed(0)
        glow = G.Glow(make_hps((16, 16, 3), K=2, L=3, hidden_channels=8, batch=2, devices=("cpu",)))
        glow.h_top.requires_grad_(False)
        arena = FlatArena(glow)
        ranges = arena_level_ranges(glow, arena)
        assert len(ranges) == 3 and ranges[0][0] == 0 and ranges[-1][1] == arena.numel
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        # every parameter lies in the slice of its own level (flow.layers: squeeze, 2 steps, split, squeeze, ...)
        for name, off in zip(arena.names, arena.offsets):
            layer = int(name.split(".")[2])
            level = 0 if layer < 4 else (1 if layer < 8 else 2)
            assert ranges[level][0] <= off < ranges[level][1], (name, off, ranges)
        g = torch.Generator().manual_seed(100 + rank)
        arena.grad.copy_(torch.randn(arena.numel, generator=g))
        whole = arena.grad.clone()
        allreduce_mean_(whole)
        works = [allreduce_mean_(arena.grad[lo:hi], async_op=True) for lo, hi in reversed(ranges)]   # top level first
        for w in works:
            w.wait()
        out[rank] = bool(torch.allclose(arena.grad, whole, rtol=0, atol=1e-7))
    finally:
        dist.destroy_process_group()


def test_per_level_allreduce_matches_whole_arena_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_levels, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and all(out[r] for r in range(world))
