"""Loaders that turn tests/golden/*.npz (outputs of the unmodified reference, see make_golden.py) into
the plain-tensor arguments of the oracle and of the CUDA layers (test helper)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def sss_lists(z):
    n = len(z["dims_in"])
    return [[torch.tensor(z[f"{name}{k}"]) for k in range(n)] for name in "ABCDEFG"]


def sss_grad_lists(z):
    n = len(z["dims_in"])
    return [[z[f"g{name}{k}"] for k in range(n)] for name in "ABCDEFG"]


def hmat_components(z):
    comps = []
    for c in range(int(z["ncomp"])):
        r0, r1, c0, c1 = [int(v) for v in z[f"rng{c}"]]
        comps.append((r0, r1, c0, c1, torch.tensor(z[f"L{c}"]), torch.tensor(z[f"R{c}"])))
    return comps


def ldr_rep(z):
    n = z["G"].shape[0]
    A = torch.sparse_coo_tensor(torch.tensor(z["A_idx"]), torch.tensor(z["A_val"]), (n, n))
    B = torch.sparse_coo_tensor(torch.tensor(z["B_idx"]), torch.tensor(z["B_val"]), (n, n))
    return [A, B, torch.tensor(z["G"]), torch.tensor(z["H"])]
