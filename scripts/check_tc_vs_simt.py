"""Full-batch consistency of the tensor-core SSS path against the SIMT path (both on the GPU): outputs and the flat gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sysm = random_mixed_system(4096, 1000, 500, 16, seed=5000)
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm).to("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
gy = (torch.rand((B, 1000), device="cuda", generator=g) * 2 - 1) / B
res = {}
for mode in ("tc", "simt"):
    os.environ["SNB200_SSS_PATH"] = mode
    layer.zero_flat_grad()
    for p in layer.parameters():
        p.grad = None
    y = layer(x)
    y.backward(gy)
    torch.cuda.synchronize()
    res[mode] = (y.detach().double(), layer.flat_grad().detach().double().clone())
ey = float((res["tc"][0] - res["simt"][0]).abs().max() / res["simt"][0].abs().max())
eg = float((res["tc"][1] - res["simt"][1]).abs().max() / res["simt"][1].abs().max())
print("B", B, "rel err y", ey, "rel err flat grad", eg)
