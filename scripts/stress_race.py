"""Determinism stress for the tensor-core SSS path: the forward has no atomics, so repeated runs must be bit-identical; the gradients
may differ by summation order only (~1e-6).  Prints the first deviations with the region of the flat gradient they fall in."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structurednets_b200.layers.sss_layer import SSSLayer  # noqa: E402
from structurednets_b200.synth import random_mixed_system  # noqa: E402


def region(layer, idx):
    off = 0
    for name, p in layer.named_parameters():
        n = p.numel()
        if off <= idx < off + n:
            return name
        off += n
    return "?"


def poison():
    """Fills the caching allocator's free blocks with NaN so that a kernel reading workspace it never wrote shows up deterministically."""
    ts = [torch.full((1 << 28,), float("nan"), device="cuda") for _ in range(3)] + [torch.full((1 << 26,), float("nan"), device="cuda") for _ in range(8)]
    torch.cuda.synchronize()
    del ts


def run(i, o, n, d, B, reps, seed):
    os.environ["SNB200_SSS_PATH"] = "tc"
    layer = SSSLayer(i, o, 0.9 if i < 1000 else 0.105, nb_states=n, initial_system_approx=random_mixed_system(i, o, n, d, seed=seed)).to("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((B, i), device="cuda", generator=g) * 2 - 1
    gy = (torch.rand((B, o), device="cuda", generator=g) * 2 - 1) / B
    ref = None
    bad = 0
    for r in range(reps):
        layer.zero_flat_grad()
        for p in layer.parameters():
            p.grad = None
        if os.environ.get("SNB200_STRESS_POISON"):
            poison()
        y = layer(x)
        y.backward(gy)
        torch.cuda.synchronize()
        cur = (y.detach().clone(), layer.flat_grad().detach().clone())
        if not bool(torch.isfinite(cur[0]).all()) or not bool(torch.isfinite(cur[1]).all()):
            nf = (~torch.isfinite(cur[1])).nonzero()
            print("  rep %d: non-finite values: y %d, grad %d (first grad index %s: %s)" % (
                r, int((~torch.isfinite(cur[0])).sum()), int(nf.numel()), int(nf[0]) if nf.numel() else -1,
                region(layer, int(nf[0])) if nf.numel() else "-"))
        if ref is None:
            ref = cur
            continue
        dy = (cur[0] - ref[0]).abs()
        dg = (cur[1] - ref[1]).abs()
        ey = float(dy.max() / ref[0].abs().max())
        eg = float(dg.max() / ref[1].abs().max())
        if ey > 0 or eg > 1e-5:
            bad += 1
            iy = int(dy.argmax())
            ig = int(dg.argmax())
            print("  rep %d: y dev %.3e at (row %d, col %d); grad dev %.3e at %d (%s); rows with y dev: %d" % (
                r, ey, iy // o, iy % o, eg, ig, region(layer, ig), int((dy.max(dim=1).values > 0).sum())))
    print("case %dx%d n=%d B=%d: %d / %d repetitions deviate" % (i, o, n, B, bad, reps - 1))


if __name__ == "__main__":
    for chain in ("1", "0"):
        os.environ["SNB200_SSS_TC_CHAIN"] = chain
        print("SNB200_SSS_TC_CHAIN=" + chain)
        if not os.environ.get("SNB200_STRESS_POISON"):
            run(320, 64, 40, 16, 150, 40, 0)
            run(128, 24, 12, 16, 33, 40, 0)
        run(4096, 1000, 500, 16, 65536 if chain == "1" else 8192, 6, 5000)
