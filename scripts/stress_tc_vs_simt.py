"""Stress for test_full_batch_gradients_tc_vs_simt: alternate the tensor-core and the SIMT path on one layer, with the caching
allocator's free blocks poisoned with NaNs before every pass (an uninitialised read then shows up as NaN / a large error)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
gy = (torch.rand((B, 1000), device="cuda", generator=g) * 2 - 1) / B


def poison():
    junk = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 28, 1 << 27, 1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18)]
    del junk


first = {}
for r in range(reps):
    for mode in ("tc", "simt"):
        os.environ["SNB200_SSS_PATH"] = mode
        poison()
        layer.zero_flat_grad()
        for p in layer.parameters():
            p.grad = None
        y = layer(x)
        y.backward(gy)
        torch.cuda.synchronize()
        cur = (y.detach().double(), layer.flat_grad().detach().double().clone())
        del y
        if mode not in first:
            first[mode] = cur
        ey = float((cur[0] - first[mode][0]).abs().max() / first[mode][0].abs().max())
        eg = float((cur[1] - first[mode][1]).abs().max() / first[mode][1].abs().max())
        nan = bool(torch.isnan(cur[1]).any() or torch.isnan(cur[0]).any())
        if ey > 1e-5 or eg > 1e-5 or nan:
            print("rep %d mode %s: vs first pass of the mode  y %.3e  grad %.3e  nan %s" % (r, mode, ey, eg, nan))
ey = float((first["tc"][0] - first["simt"][0]).abs().max() / first["simt"][0].abs().max())
eg = float((first["tc"][1] - first["simt"][1]).abs().max() / first["simt"][1].abs().max())
print("tc vs simt (first passes): y %.3e grad %.3e" % (ey, eg))
