"""Data-parallel plumbing on CPU: world_size 2, gloo (the CUDA kernels are not involved; gradients are synthetic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from structurednets_b200.distributed import GradSynchronizer, broadcast_parameters, shard_bounds
from structurednets_b200.layers.lr_layer import LRLayer
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)   # different initial parameters per rank on purpose
        np.random.seed(rank)
        sss = SSSLayer(24, 8, 0.9, nb_states=4, initial_system_approx=random_mixed_system(24, 8, 4, 2, seed=10 + rank))
        lr = LRLayer(24, 8, 0.5)
        model = torch.nn.ModuleList([sss, lr])
        broadcast_parameters(model, src=0)
        ref = torch.cat([sss.flat_parameters().clone(), lr.flat_parameters().clone()])
        # synthetic per-rank gradients written where the backward kernels would put them
        for m in (sss, lr):
            g = m._prepare_grad_accumulation()
            g.copy_(torch.arange(g.numel(), dtype=torch.float32) * (rank + 1))
        GradSynchronizer(model, scale=0.5)()
        total = sum(r + 1 for r in range(world))
        ok = True
        for m in (sss, lr):
            g = m.flat_grad()
            ok &= bool(torch.allclose(g, torch.arange(g.numel(), dtype=torch.float32) * total * 0.5))
            ok &= all(p.grad is not None and p.grad.data_ptr() >= g.data_ptr() for p in m.parameters() if p.numel())
        gathered = [torch.zeros_like(ref) for _ in range(world)]
        dist.all_gather(gathered, ref)
        ok &= all(bool(torch.equal(t, gathered[0])) for t in gathered)
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_and_broadcast_world_size_2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_bounds_partition_the_batch():
    for n, w in ((65536, 8), (10, 3), (7, 8)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
