"""Bitwise determinism of the forward's saved states S[chunk][B][32] (and y) over repeated sn_sss_tc_forward calls at C5's batch."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structurednets_b200 import _lib  # noqa: E402
from structurednets_b200.layers.sss_layer import SSSLayer  # noqa: E402
from structurednets_b200.synth import random_mixed_system  # noqa: E402

B = int(os.environ.get("STRESS_B", "65536"))
reps = int(os.environ.get("STRESS_REPS", "40"))
os.environ["SNB200_SSS_PATH"] = "tc"
layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
L = _lib.lib()
layer._ensure_flat()
tc = layer._tc_plan(dev)
ps = ctypes.byref(tc["struct"])
nc = tc["struct"].nchunks
flat = layer.flat_parameters()
_lib.check(L.sn_sss_tc_build(ps, _lib.ptr(flat), _lib.ptr(tc["coef"]), _lib.stream_ptr()), "build")
ref = None
bad = 0
for r in range(reps):
    y = torch.empty((B, 1000), device=dev)
    rbuf = torch.full((int(L.sn_sss_tc_rbuf_floats(ps, B)),), float("nan"), device=dev)
    states = torch.full((int(L.sn_sss_tc_states_floats(ps, B)),), float("nan"), device=dev)
    _lib.check(L.sn_sss_tc_forward(ps, _lib.ptr(tc["coef"]), _lib.ptr(x), x.stride(0), _lib.ptr(y), y.stride(0), _lib.ptr(layer.bias),
                                   _lib.ptr(rbuf), _lib.ptr(states), B, _lib.stream_ptr()), "forward")
    torch.cuda.synchronize()
    S = states.view(nc, B, 32)
    if ref is None:
        ref = (y.clone(), S.clone())
        print("non-finite in S:", int((~torch.isfinite(S)).sum()), " in y:", int((~torch.isfinite(y)).sum()))
        continue
    dS = (S != ref[1]) & ~(torch.isnan(S) & torch.isnan(ref[1]))
    dy = y != ref[0]
    if bool(dS.any()) or bool(dy.any()):
        bad += 1
        idx = dS.nonzero()
        ch = sorted(set(idx[:, 0].tolist()))
        rows = idx[:, 1]
        halves = sorted(set((idx[:, 2] // 16).tolist()))
        print("rep %d: S differs in %d entries: chunks %s, rows %d..%d (tiles %s), halves %s; y differs in %d entries" % (
            r, int(dS.sum()), ch[:8], int(rows.min()) if len(rows) else -1, int(rows.max()) if len(rows) else -1,
            sorted(set((rows // 128).tolist()))[:8], halves, int(dy.sum())))
        if len(idx):
            c0, r0, k0 = [int(v) for v in idx[0]]
            print("   S[%d][%d][%d..]: got %s | ref %s" % (c0, r0, k0, S[c0, r0, 16:32].tolist(), ref[1][c0, r0, 16:32].tolist()))
        if bool(dy.any()):
            iy = dy.nonzero()[0]
            r1, c1 = int(iy[0]), int(iy[1])
            print("   y[%d][%d..]: got %s | ref %s" % (r1, c1, y[r1, c1:c1 + 8].tolist(), ref[0][r1, c1:c1 + 8].tolist()))
print("B=%d: %d / %d repetitions deviate" % (B, bad, reps - 1))
