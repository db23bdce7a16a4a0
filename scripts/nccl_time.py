"""torchrun --nproc-per-node N scripts/nccl_time.py: time of torch.distributed.all_reduce on the flat gradient of the SSS C5 layer
(430 503 floats) inside a CUDA graph, for the NCCL settings of the environment (NCCL_ALGO / NCCL_PROTO ...)."""
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
n = 430503
buf = torch.randn(n, device=dev)
for _ in range(5):
    dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
t = torch.tensor([best], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world %d NCCL_ALGO=%s NCCL_PROTO=%s NCCL_NVLS_ENABLE=%s: all_reduce %.1f us" % (world, os.environ.get("NCCL_ALGO"), os.environ.get("NCCL_PROTO"),
                                                                                         os.environ.get("NCCL_NVLS_ENABLE"), float(t.item())), flush=True)
del g
dist.barrier()
dist.destroy_process_group()
