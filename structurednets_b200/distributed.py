"""Data-parallel training of a structured layer across the GPUs of one box (BASELINE config C5).

The reference has no distributed code.  The path shards naturally along the batch only:
every rank holds the full (tiny) parameter set and a slice of the feature batch, runs the identical fused
forward/backward on its slice, and the gradients are summed with ONE collective per step over the layer's
single flat gradient buffer (``FlatParamsMixin.flat_grad()``; sparse parameters contribute their value
arrays, the pattern being identical on every rank).  ``torch.distributed`` is plumbing: ``nccl`` on GPUs
(NVLink 5 / NVSwitch), ``gloo`` in the CPU tests.
"""
import os
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


class P2PAllReduce:
    """One-shot all-reduce (sum) of a small fp32 CUDA buffer over NVLink peer memory: ``csrc/p2p.cu`` (``sn_allreduce_oneshot_f32``).

    Every rank allocates one region through the library, the 64-byte CUDA IPC handles are exchanged with ``all_gather_object`` and every
    rank maps its peers' regions.  ``__call__(buf, scale)`` then costs one kernel launch per step (push the local slice into every rank's
    region, one flag barrier among the blocks that own the slice, sum the slots in rank order), is bit-identical on every rank and can be
    captured in a CUDA graph.  Measured for the 1.7 MB gradient of the 500-stage SSS layer (``scripts/p2p_test.py``, 20 calls per graph):
    2 GPUs 20.5 us against 20.9 us for NCCL's all-reduce + scale, 8 GPUs 44.9 us against 35.7 us (a first, pulling version: 24.6 us on two
    GPUs).  One NVSwitch round trip for the system-scope fence plus one for the flags costs what NCCL's low-latency protocol costs in
    total, and at 8 GPUs every rank pushes 7 x 1.7 MB.  It is therefore OPT-IN (``GradSynchronizer(p2p=True)`` / SNB200_P2P_ALLREDUCE=1) and
    NCCL stays the default transport of the path's one exchange step.  Construction is collective; ``P2PAllReduce.create`` returns None
    on every rank if any rank cannot set it up (no peer access, IPC not permitted), and the caller keeps using NCCL."""

    MAX_FLOATS = 4 << 20          # 16 MB: beyond that the exchange is bandwidth bound and a ring all-reduce moves fewer bytes per link

    def __init__(self, capacity_floats: int, group=None):
        from . import _lib
        import ctypes
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(capacity_floats)
        L = _lib.lib()
        own = ctypes.c_void_p()
        _lib.check(L.sn_p2p_alloc(ctypes.byref(own), self.capacity, self.world), "p2p_alloc")
        self._own = own
        handle = ctypes.create_string_buffer(64)
        _lib.check(L.sn_p2p_export(own, handle), "p2p_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.rank, bytes(handle.raw)), group=group)
        self._regions = (ctypes.c_void_p * self.world)()
        self._imported = []
        for r, raw in handles:
            if r == self.rank:
                self._regions[r] = own.value
            else:
                peer = ctypes.c_void_p()
                _lib.check(L.sn_p2p_import(ctypes.create_string_buffer(raw, 64), ctypes.byref(peer)), "p2p_import")
                self._regions[r] = peer.value
                self._imported.append(peer)
        import atexit
        atexit.register(self.close)      # a process that exits with its peers' regions still mapped can hang in the driver's teardown

    def close(self):
        """Collective: unmaps the peers' regions and frees the own one (also registered with atexit: a process must not exit with its
        peers' memory still mapped while they are still running kernels on it)."""
        if self._own is None:
            return
        from . import _lib
        L = _lib.lib()
        torch.cuda.synchronize()
        if dist.is_initialized():
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
        for peer in self._imported:
            L.sn_p2p_close(peer)
        self._imported = []
        L.sn_p2p_free(self._own)
        self._own = None

    @classmethod
    def create(cls, capacity_floats: int, group=None, device=None):
        """Collective.  The P2P all-reduce on every rank, or None on every rank."""
        obj, ok = None, 1
        try:
            obj = cls(capacity_floats, group)
        except Exception as e:      # peer access / IPC unavailable: agree on NCCL below
            ok = 0
            import warnings
            warnings.warn("P2P all-reduce unavailable on rank %d (%s): using NCCL" % (dist.get_rank(group), e))
        flag = torch.tensor([ok], dtype=torch.int32, device=device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return obj if int(flag.item()) == 1 else None

    def __call__(self, buf: torch.Tensor, scale: float = None):
        from . import _lib
        assert buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous() and buf.numel() <= self.capacity
        _lib.check(_lib.lib().sn_allreduce_oneshot_f32(_lib.ptr(buf), buf.numel(), self._regions, self.rank, self.world, self.capacity,
                                                       1.0 if scale is None else float(scale), _lib.stream_ptr()), "allreduce_oneshot")


class GradSynchronizer:
    """Callable that all-reduces (sum) the gradients of `module` over `group` and optionally rescales them.

    ``scale``: factor applied after the sum; use ``local_batch / global_batch`` when each rank's loss is a mean
    over its local shard and the reference semantics (mean over the global batch) are wanted.

    The flat gradient buffers and the list of parameters living outside them are looked up once (walking the 3 501
    parameter views of an SSS layer on every step costs more host time than the step's kernels leave at 8 GPUs).
    """

    def __init__(self, module: torch.nn.Module, group=None, scale: float = None, p2p: bool = None):
        self.module, self.group, self.scale = module, group, scale
        self._flat_modules = None
        self._loose = None
        # opt-in (p2p=True or SNB200_P2P_ALLREDUCE=1): fp32 CUDA buffers of up to 16 MB go through the one-shot NVLink all-reduce of
        # csrc/p2p.cu, set up collectively on the first call.  NCCL is the default because it is faster on this box (P2PAllReduce docstring).
        self._p2p_wanted = (os.environ.get("SNB200_P2P_ALLREDUCE", "0") == "1") if p2p is None else bool(p2p)
        self._p2p = None
        self._p2p_tried = False

    def _prepare(self):
        self._flat_modules, covered = [], set()
        for m in self.module.modules():
            if hasattr(m, "flat_grad") and hasattr(m, "_flat_param_list"):
                self._flat_modules.append(m)
                covered.update(id(p) for p in m._flat_param_list())
        self._loose = [p for p in self.module.parameters() if id(p) not in covered]

    @property
    def transport(self) -> str:
        """'p2p' once the one-shot NVLink all-reduce is set up, else 'nccl' / the group's backend."""
        return "p2p" if self._p2p is not None else (dist.get_backend(self.group) if dist.is_initialized() else "none")

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
        bufs = self.buffers()
        if self._p2p_wanted and not self._p2p_tried and not (bufs and bufs[0].is_cuda and torch.cuda.is_current_stream_capturing()):
            self._p2p_tried = True      # (the set-up allocates and synchronises: never inside a graph capture -- call prepare() before)
            eligible = [b.numel() for b in bufs if b.is_cuda and b.dtype == torch.float32 and b.is_contiguous() and b.numel() <= P2PAllReduce.MAX_FLOATS]
            want = torch.tensor([max(eligible) if eligible else 0], dtype=torch.int64, device=bufs[0].device if bufs else None)
            if bufs and bufs[0].is_cuda and dist.get_backend(self.group) == "nccl":
                dist.all_reduce(want, op=dist.ReduceOp.MAX, group=self.group)      # same capacity on every rank
                if int(want.item()) > 0:
                    self._p2p = P2PAllReduce.create(int(want.item()), self.group, device=bufs[0].device)
        for buf in bufs:
            if (self._p2p is not None and buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous() and buf.numel() <= self._p2p.capacity
                    and buf.data_ptr() % 16 == 0):
                self._p2p(buf, self.scale)
                continue
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
