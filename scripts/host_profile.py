"""Host-side cost of one eager SSS step (forward + backward through the module) at a small batch: cProfile of 300 steps."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
x = torch.rand((B, 4096), device="cuda")
gy = torch.rand((B, 1000), device="cuda") / B


def step():
    layer.zero_flat_grad()
    y = layer(x)
    y.backward(gy)


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("B=%d: host enqueue %.1f us per step, with the device drained %.1f us per step" % (B, t_host / 300 * 1e6, t_all / 300 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
