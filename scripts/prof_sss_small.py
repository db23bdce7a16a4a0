"""Profiling driver: a few forward + backward passes of the C1 SSS layer at a small batch (the batch-independent kernels dominate).
    ncu --set full --import-source on -k regex:build -o gpurun_out/prof python scripts/prof_sss_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
gy = (torch.rand((B, 1000), device="cuda", generator=g) * 2 - 1) / B
for _ in range(reps):
    layer.zero_flat_grad()
    y = layer(x)
    y.backward(gy)
torch.cuda.synchronize()
print("done")
