"""torchrun --nproc-per-node N scripts/p2p_test.py: the one-shot NVLink all-reduce against NCCL (values, bit-identity across ranks, graph
capture, timing)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from structurednets_b200.distributed import P2PAllReduce

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
n = 430503
log = open("gpurun_out/p2p_rank%d.log" % rank, "w")
def say(*a):
    print(*a, file=log, flush=True)
say("process group up")
ar = P2PAllReduce.create(n, device=dev)
say("created", ar is not None)
assert ar is not None, "P2P all-reduce could not be set up"
ok = True
for rep in range(20):
    g = torch.Generator(device=dev).manual_seed(100 * rep + rank)
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    y = x.clone()
    torch.cuda.synchronize()
    say("rep", rep, "launching")
    ar(y, 0.5)
    torch.cuda.synchronize()
    say("rep", rep, "done")
    err = float((y - 0.5 * ref).abs().max() / ref.abs().max())
    gathered = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(gathered, y)
    same = all(bool(torch.equal(t, gathered[0])) for t in gathered)
    ok = ok and err < 1e-6 and same
    if rank == 0 and rep < 3:
        print("rep %d: max rel err vs NCCL %.2e, identical on all ranks %s" % (rep, err, same), flush=True)
# graph capture + timing
buf = torch.randn(n, device=dev)
for _ in range(3):
    ar(buf)
torch.cuda.synchronize(); dist.barrier()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(20):
        ar(buf, 1.0 / world)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
graph.replay(); torch.cuda.synchronize(); dist.barrier()
e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
t_p2p = e0.elapsed_time(e1) / 20 * 1e3
buf2 = torch.randn(n, device=dev)
g2 = torch.cuda.CUDAGraph()
for _ in range(3):
    dist.all_reduce(buf2)
torch.cuda.synchronize(); dist.barrier()
with torch.cuda.graph(g2):
    for _ in range(20):
        dist.all_reduce(buf2)
        buf2.mul_(1.0 / world)
torch.cuda.synchronize(); dist.barrier()
g2.replay(); torch.cuda.synchronize(); dist.barrier()
e0.record(); g2.replay(); e1.record(); torch.cuda.synchronize()
t_nccl = e0.elapsed_time(e1) / 20 * 1e3
if rank == 0:
    print("world %d, %d floats: one-shot P2P %.1f us per all-reduce, NCCL all_reduce + scale %.1f us; all checks %s" % (world, n, t_p2p, t_nccl, "passed" if ok else "FAILED"), flush=True)
del graph, g2
say("closing")
ar.close()
say("closed")
dist.barrier()
dist.destroy_process_group()
say("destroyed")
sys.exit(0 if ok else 1)
