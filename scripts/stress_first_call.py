"""One fresh process: TC fwd+bwd twice, SIMT fwd+bwd twice at C5's batch; reports any deviation between the four gradient vectors
and the region of the flat gradient it falls in.  Run many times (the intermittent mismatch of test_full_batch_gradients_tc_vs_simt
shows up about once in twelve processes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structurednets_b200.layers.sss_layer import SSSLayer  # noqa: E402
from structurednets_b200.synth import random_mixed_system  # noqa: E402


def region(layer, idx):
    off = 0
    for name, p in layer.named_parameters():
        n = p.numel()
        if off <= idx < off + n:
            return "%s[%d]" % (name, idx - off)
        off += n
    return "?"


B = 65536
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
gy = (torch.rand((B, 1000), device="cuda", generator=g) * 2 - 1) / B
res = []
for mode in ("tc", "simt", "tc", "simt"):
    os.environ["SNB200_SSS_PATH"] = mode
    layer.zero_flat_grad()
    for p in layer.parameters():
        p.grad = None
    y = layer(x)
    y.backward(gy)
    torch.cuda.synchronize()
    res.append((mode, layer.flat_grad().detach().double().clone()))
    del y
scale = float(res[1][1].abs().max())
msgs = []
for a in range(4):
    for b in range(a + 1, 4):
        d = (res[a][1] - res[b][1]).abs()
        e = float(d.max()) / scale
        if e > 1e-5:
            i = int(d.argmax())
            msgs.append("%s#%d vs %s#%d: %.2e at %s, entries above 1e-5: %d" % (res[a][0], a, res[b][0], b, e, region(layer, i), int((d / scale > 1e-5).sum())))
print("OK" if not msgs else "DEVIATION " + " | ".join(msgs))
