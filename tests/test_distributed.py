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
        # p2p=True asks for the one-shot NVLink all-reduce; on CPU tensors / gloo it must fall back to the group's all_reduce
        sync = GradSynchronizer(model, scale=0.5, p2p=True)
        sync()
        assert sync.transport == "gloo"
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


def _train_worker(rank, world, port, out):
    """Data-parallel train_resident: every rank trains on its shard of each batch with one gradient all-reduce per step; the
    parameters must stay identical across ranks and equal single-process training on the whole batches."""
    from structurednets_b200 import training_helpers as TH
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        X = rng.uniform(-1, 1, size=(96, 10)).astype(np.float32)
        Y = rng.uniform(-1, 1, size=(96, 3)).astype(np.float32)
        # shard every batch of 32 rows: rank r takes rows [16 r, 16 r + 16) of each batch; an identity "shuffle" keeps the batches aligned
        idx = np.concatenate([np.arange(b * 32 + rank * 16, b * 32 + rank * 16 + 16) for b in range(3)])
        torch.manual_seed(0)
        model = torch.nn.Linear(10, 3)
        orig_shuffle = TH.shuffle
        TH.shuffle = lambda *a: a[0] if len(a) == 1 else a
        try:
            res = TH.train_resident(model, X[idx], Y[idx], X_val=X[:8], y_val=Y[:8], patience=1, batch_size=16, lr=0.05,
                                    loss_function_class=torch.nn.MSELoss, min_patience_improvement=1e6, optimizer_class=torch.optim.SGD,
                                    restore_best_model=False, grad_sync=GradSynchronizer(model, scale=1.0 / world))
            params = torch.cat([p.detach().reshape(-1) for p in res[0].parameters()])
            gathered = [torch.zeros_like(params) for _ in range(world)]
            dist.all_gather(gathered, params)
            same = all(bool(torch.equal(g, gathered[0])) for g in gathered)
            # single-process reference on the whole batches of 32
            torch.manual_seed(0)
            ref_model = torch.nn.Linear(10, 3)
            ref = TH.train_resident(ref_model, X, Y, X_val=X[:8], y_val=Y[:8], patience=1, batch_size=32, lr=0.05,
                                    loss_function_class=torch.nn.MSELoss, min_patience_improvement=1e6, optimizer_class=torch.optim.SGD,
                                    restore_best_model=False)
            ref_params = torch.cat([p.detach().reshape(-1) for p in ref[0].parameters()])
            out[rank] = same and bool(torch.allclose(params, ref_params, rtol=1e-5, atol=1e-6)) and len(res[5]) == len(ref[5]) == 2
        finally:
            TH.shuffle = orig_shuffle
    finally:
        dist.destroy_process_group()


def test_data_parallel_train_resident_world_size_2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_train_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
