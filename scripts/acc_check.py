"""Accuracy of the tensor-core SSS path against the oracle in float64 at the C1 shape: max-abs-relative error of y and of every parameter
list's gradient (tests use 1e-5 against the float32 oracle; this prints the actual margins).  Usage: python scripts/acc_check.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=1001))
rng = np.random.default_rng(1001)
X = rng.uniform(-1, 1, size=(B, 4096)).astype(np.float32)
gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
l64 = [[p.detach().double().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]
b = layer.bias.detach().double().clone().requires_grad_(True)
yo = O.sss_forward(torch.tensor(X).double(), *l64, b, layer.dims_in, layer.dims_out)
(yo * torch.tensor(gy).double()).sum().backward()
layer = layer.to("cuda")
Xd, gyd = torch.tensor(X, device="cuda"), torch.tensor(gy, device="cuda")


def rel(a, r):
    return float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-30))


for mode in ("tc", "simt"):
    os.environ["SNB200_SSS_PATH"] = mode
    for p in layer.parameters():
        p.grad = None
    y = layer(Xd)
    (y * gyd).sum().backward()
    torch.cuda.synchronize()
    errs = {"y": rel(y.detach().cpu().numpy().astype(np.float64), yo.detach().numpy())}
    for li, name in enumerate("ABCDEFG"):
        got = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for p in getattr(layer, name)]).astype(np.float64)
        ref = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape))).reshape(-1) for p in l64[li]])
        if ref.size:
            errs["d" + name] = rel(got, ref)
    errs["dbias"] = rel(layer.bias.grad.detach().cpu().numpy().astype(np.float64), b.grad.numpy())
    print(mode, " ".join("%s %.2e" % kv for kv in errs.items()))
