"""CPU suite, part 4: the N>1 host logic under gloo, world_size 2 (no GPU involved)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_decode_b200 import codes
from gnn_decode_b200.dist import allreduce_counts, allreduce_flat_grads, shard_bounds
from oracle import philox


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 15, 16, 1000, 65536, 65541):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            if n >= 8 * world:
                assert all(s % 8 == 0 for s, _ in spans)
                sizes = [e - s for s, e in spans]
                assert max(sizes) - min(sizes) <= 8 + n % 8
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) sharded sampling == slices of the single-process draw (no collective on the data path)
        pcm = codes.toric_pcm(4)
        B = 100
        s, e = shard_bounds(B, rank, world)
        x, err = philox.sample(pcm, e - s, [0.05, 0.1], noise=0, seed=77, first_sample=s)
        # (2) flat gradient all-reduce
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(2, 4), torch.nn.Linear(4, 1))
        for i, p in enumerate(lin.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        flat = allreduce_flat_grads(lin.parameters(), average=False)
        cnt = allreduce_counts(torch.tensor([rank + 1, 10 * (rank + 1), 0]))
        out[rank] = (s, e, x, err, [p.grad.clone() for p in lin.parameters()], flat.numel(), cnt)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    pcm = codes.toric_pcm(4)
    x_full, e_full = philox.sample(pcm, 100, [0.05, 0.1], noise=0, seed=77, first_sample=0)
    for r in range(world):
        s, e, x, err, grads, n_flat, cnt = out[r]
        assert np.array_equal(x, x_full[s:e]) and np.array_equal(err, e_full[s:e])
        for i, g in enumerate(grads):
            assert torch.all(g == 3.0 * (i + 1))        # (1 + 2) * (i + 1)
        assert n_flat == 2 * 4 + 4 + 4 + 1
        assert cnt.tolist() == [3, 30, 0]
