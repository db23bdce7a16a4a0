"""Stress: repeated SSS forward + backward on the tensor-core path; the flat gradient of every repetition must agree with the first
one (to atomics' summation-order noise).  SNB200_SSS_BUILD=col|quad|(default) selects the build kernels.
    python scripts/stress_grad.py [batch] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
os.environ.setdefault("SNB200_SSS_PATH", "tc")
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
gy = (torch.rand((B, 1000), device="cuda", generator=g) * 2 - 1) / B
ref = None
bad = 0
offs = layer._param_offsets()
for r in range(reps):
    layer.zero_flat_grad()
    y = layer(x)
    y.backward(gy)
    torch.cuda.synchronize()
    gcur = layer.flat_grad().detach().double().clone()
    if ref is None:
        ref = gcur
        continue
    err = float((gcur - ref).abs().max() / ref.abs().max())
    if err > 1e-5:
        bad += 1
        idx = int((gcur - ref).abs().argmax())
        which = max(((k, o) for k, o in offs.items() if o <= idx), key=lambda t: t[1])
        print("rep %d: rel err %.3e at flat index %d (%s.%d + %d): %g vs %g; entries off by > 1e-5: %d" % (
            r, err, idx, which[0][0], which[0][1], idx - which[1], float(gcur[idx]), float(ref[idx]),
            int(((gcur - ref).abs() > 1e-5 * ref.abs().max()).sum())))
print("build=%s batch=%d: %d deviating repetitions of %d" % (os.environ.get("SNB200_SSS_BUILD", "mma"), B, bad, reps - 1))
