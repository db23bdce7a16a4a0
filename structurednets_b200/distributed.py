"""Data-parallel training of a structured layer across the GPUs of one box (BASELINE config C5).

The reference has no distributed code.  The path shards naturally along the batch only:
every rank holds the full (tiny) parameter set and a slice of the feature batch, runs the identical fused
forward/backward on its slice, and the gradients are summed with ONE collective per step over the layer's
single flat gradient buffer (``FlatParamsMixin.flat_grad()``; sparse parameters contribute their value
arrays, the pattern being identical on every rank).  ``torch.distributed`` is plumbing: ``nccl`` on GPUs
(NVLink 5 / NVSwitch), ``gloo`` in the CPU tests.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_bounds(n_samples: int, rank: int, world_size: int):
    """Contiguous, balanced row shard [lo, hi) of a batch of n_samples for `rank`."""
    base, rem = divmod(n_samples, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _grad_buffers(module: torch.nn.Module) -> List[torch.Tensor]:
    """The tensors that must be summed across ranks: one flat buffer for modules with flat parameters,
    plus the gradient (value arrays for sparse ones) of every parameter living outside it."""
    bufs = []
    covered = set()
    for m in module.modules():
        if hasattr(m, "flat_grad") and hasattr(m, "_flat_param_list"):
            if m.__dict__.get("_flat_grad") is not None:
                bufs.append(m.__dict__["_flat_grad"])
            covered.update(id(p) for p in m._flat_param_list())
    for p in module.parameters():
        if id(p) in covered or p.grad is None:
            continue
        bufs.append(p.grad._values() if p.grad.is_sparse else p.grad)
    return bufs


class GradSynchronizer:
    """Callable that all-reduces (sum) the gradients of `module` over `group` and optionally rescales them.

    ``scale``: factor applied after the sum; use ``local_batch / global_batch`` when each rank's loss is a mean
    over its local shard and the reference semantics (mean over the global batch) are wanted.

    The flat gradient buffers and the list of parameters living outside them are looked up once (walking the 3 501
    parameter views of an SSS layer on every step costs more host time than the step's kernels leave at 8 GPUs).
    """

    def __init__(self, module: torch.nn.Module, group=None, scale: float = None):
        self.module, self.group, self.scale = module, group, scale
        self._flat_modules = None
        self._loose = None

    def _prepare(self):
        self._flat_modules, covered = [], set()
        for m in self.module.modules():
            if hasattr(m, "flat_grad") and hasattr(m, "_flat_param_list"):
                self._flat_modules.append(m)
                covered.update(id(p) for p in m._flat_param_list())
        self._loose = [p for p in self.module.parameters() if id(p) not in covered]

    def buffers(self) -> List[torch.Tensor]:
        if self._flat_modules is None:
            self._prepare()
        bufs = [m.__dict__["_flat_grad"] for m in self._flat_modules if m.__dict__.get("_flat_grad") is not None]
        for p in self._loose:
            if p.grad is not None:
                bufs.append(p.grad._values() if p.grad.is_sparse else p.grad)
        return bufs

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        for buf in self.buffers():
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            if self.scale is not None:
                buf.mul_(self.scale)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Make every rank start from rank `src`'s parameters (one broadcast per flat buffer / loose parameter)."""
    seen = set()
    for m in module.modules():
        if hasattr(m, "flat_parameters") and hasattr(m, "_flat_param_list"):
            dist.broadcast(m.flat_parameters(), src=src, group=group)
            seen.update(id(p) for p in m._flat_param_list())
    for p in module.parameters():
        if id(p) not in seen:
            dist.broadcast(p._values() if p.is_sparse else p.data, src=src, group=group)
