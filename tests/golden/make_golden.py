"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden.py

The reference ships no golden vectors.  This script imports the reference from
/root/reference/src (with the four absent third-party imports stubbed, oracle/ref_import.py), builds
each layer through its own constructor from seeded structures, runs forward and
``(y * gy).sum().backward()`` on CPU, and stores inputs, parameters, outputs and parameter gradients.
``tests/test_oracle_golden.py`` pins the oracle (oracle/layers_cpu.py) against these files; the GPU
parity tests then compare the CUDA path with the oracle and with these files directly.
"""
import os
import sys

import numpy as np
import scipy.sparse
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle.ref_import import load_reference  # noqa: E402
from structurednets_b200.synth import random_mixed_system  # noqa: E402


def grads_of(params):
    out = []
    for p in params:
        g = p.grad
        if g is None:
            out.append(np.zeros(tuple(p.shape), dtype=np.float64 if p.dtype == torch.float64 else np.float32))
        elif g.is_sparse:
            out.append(g.to_dense().numpy())
        else:
            out.append(g.numpy())
    return out


def run(layer, X, gy):
    y = layer(torch.tensor(X))
    (y * torch.tensor(gy)).sum().backward()
    return y.detach().numpy()


def main():
    sn = load_reference()
    rng = np.random.default_rng(20221018)

    # ---- SSS: the shape of the reference's own test_sss_layer (tests/test_layers.py:106-121) ----
    from structurednets.layers.sss_layer import SSSLayer
    for tag, (i, o, n, d, B, ragged) in {"sss_76x14": (76, 14, 10, 3, 51, False), "sss_50x50": (50, 50, 5, 4, 10, True)}.items():
        sysm = random_mixed_system(i, o, n, d, seed=42, ragged_state_dims=ragged)
        bias = rng.uniform(-1, 1, size=(o,))
        layer = SSSLayer(i, o, 0.9, initial_bias=bias, nb_states=n, initial_system_approx=sysm)
        X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32)
        gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32)
        y = run(layer, X, gy)
        # reference's own known-answer: layer == x T^T + bias
        assert np.allclose(y, X @ layer.initial_weight_matrix.T + bias, atol=1e-5)
        out = dict(X=X, gy=gy, y=y, bias=layer.bias.detach().numpy(), gbias=layer.bias.grad.numpy(),
                   dims_in=np.asarray(layer.dims_in), dims_out=np.asarray(layer.dims_out), seed=42, d=d, ragged=int(ragged))
        for name in "ABCDEFG":
            plist = getattr(layer, name)
            gs = grads_of(plist)
            for k, p in enumerate(plist):
                out[f"{name}{k}"] = p.detach().numpy()
                out[f"g{name}{k}"] = gs[k]
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)

    # ---- LR (layers/lr_layer.py) ----
    from structurednets.layers.lr_layer import LRLayer
    i, o, r, B = 96, 40, 8, 33
    L = rng.uniform(-1, 1, size=(o, r)); R = rng.uniform(-1, 1, size=(r, i)) / np.sqrt(i)
    bias = rng.uniform(-1, 1, size=(o,))
    layer = LRLayer(i, o, 0.5, initial_bias=bias, initial_lr_components=[L, R])
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32)
    y = run(layer, X, gy)
    np.savez_compressed(os.path.join(HERE, "lr_96x40.npz"), X=X, gy=gy, y=y, left=layer.left_lr.detach().numpy(),
                        right=layer.right_lr.detach().numpy(), bias=layer.bias.detach().numpy(),
                        gleft=layer.left_lr.grad.numpy(), gright=layer.right_lr.grad.numpy(), gbias=layer.bias.grad.numpy())

    # ---- PSM, 2 factors: the literal reference forward (layers/psm_layer.py:47-60) ----
    from structurednets.layers.psm_layer import PSMLayer
    i, o, B = 50, 30, 17
    mx = max(i, o)
    f0 = scipy.sparse.random(o, mx, density=0.2, random_state=1, format="csr", dtype=np.float64)
    f1 = scipy.sparse.random(mx, i, density=0.2, random_state=2, format="csr", dtype=np.float64)
    bias = rng.uniform(-1, 1, size=(o,))
    layer = PSMLayer(i, o, initial_bias=bias, sparse_matrices=[f0, f1])
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32)
    y = run(layer, X, gy)
    assert np.allclose(y, X @ (f0 @ f1).toarray().T.astype(np.float32) + bias, atol=1e-5)
    out = dict(X=X, gy=gy, y=y, bias=layer.bias.detach().numpy(), gbias=layer.bias.grad.numpy())
    for k, (f, p) in enumerate(zip([f0, f1], layer.sparse_matrices)):
        out[f"f{k}_dense"] = f.toarray()
        assert p.grad.is_sparse
        out[f"gf{k}_dense"] = p.grad.to_dense().numpy()
    np.savez_compressed(os.path.join(HERE, "psm_50x30_2f.npz"), **out)

    # ---- H-matrix (layers/hmat_layer.py) on the reference's own block cluster tree ----
    from structurednets.approximators.hmat_approximator import build_hmat_block_cluster_tree
    from structurednets.hmatrix.hmatrix import HMatrix
    from structurednets.layers.hmat_layer import HMatLayer
    o, i, B = 40, 64, 19
    tree = build_hmat_block_cluster_tree((o, i), eta=0.5, min_block_size=2)
    comps = []
    for leaf in tree.get_all_leaf_elements():
        rows, cols = len(leaf.row_range), len(leaf.col_range)
        k = int(min(3, min(rows, cols)))
        if (leaf.row_range.start + leaf.col_range.start) % 5 == 0:
            continue   # leave some leaves empty: rank-0 leaves are dropped by the layer (hmat_layer.py:26-27)
        Lk = torch.tensor(rng.uniform(-1, 1, size=(rows, k)) / np.sqrt(k)).float()
        Rk = torch.tensor(rng.uniform(-1, 1, size=(k, cols)) / np.sqrt(cols)).float()
        leaf.set_hmatrix_component(Lk, Rk)
    hm = HMatrix(block_cluster_tree=tree, shape=(o, i))
    bias = rng.uniform(-1, 1, size=(o,))
    layer = HMatLayer(i, o, 0.9, initial_bias=bias, initial_hmatrix=hm)
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32)
    y = run(layer, X, gy)
    assert np.allclose(y, X @ layer.hmatrix.to_dense_numpy().T + bias, atol=1e-5)
    out = dict(X=X, gy=gy, y=y, bias=layer.bias.detach().numpy(), gbias=layer.bias.grad.numpy(), ncomp=len(layer.hmatrix_components),
               leaf_ranges=np.asarray([[a.start, a.stop, b.start, b.stop] for (a, b) in tree.get_all_leaf_ranges()]))
    for c_i, c in enumerate(layer.hmatrix_components):
        out[f"rng{c_i}"] = np.asarray([c.row_range.start, c.row_range.stop, c.col_range.start, c.col_range.stop])
        out[f"L{c_i}"] = c.left_lr.detach().numpy(); out[f"R{c_i}"] = c.right_lr.detach().numpy()
        out[f"gL{c_i}"] = c.left_lr.grad.numpy(); out[f"gR{c_i}"] = c.right_lr.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "hmat_40x64.npz"), **out)

    # ---- LDR: the literal matrix_power construction (approximators/ldr_approximator.py:29-39) ----
    from structurednets.approximators.ldr_approximator import init_representation_matrices_torch
    from structurednets.layers.ldr_layer import LDRLayer
    n, B = 12, 9
    np.random.seed(7)
    layer = LDRLayer(n, n, 0.95)   # reference random init (np.random, seeded above); fp64 parameters
    r = layer.representation_matrices[2].shape[1]
    with torch.no_grad():   # make A, B large enough that all n Krylov columns matter
        layer.representation_matrices[0].mul_(2.0)
        layer.representation_matrices[1].mul_(2.0)
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32)
    y = run(layer, X, gy)
    A, Bm, G, H = layer.representation_matrices
    out = dict(X=X, gy=gy, y=y, bias=layer.bias.detach().numpy(), gbias=layer.bias.grad.numpy(), r=r,
               A_idx=A.detach()._indices().numpy(), A_val=A.detach()._values().numpy(),
               B_idx=Bm.detach()._indices().numpy(), B_val=Bm.detach()._values().numpy(),
               G=G.detach().numpy(), H=H.detach().numpy(),
               gA_dense=A.grad.to_dense().numpy(), gB_dense=Bm.grad.to_dense().numpy(), gG=G.grad.numpy(), gH=H.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "ldr_12.npz"), **out)

    # ---- Toeplitz-like (layers/tl_layer.py) ----
    from structurednets.layers.tl_layer import TLLayer
    n, r, B = 16, 3, 11
    G = rng.uniform(-1, 1, size=(n, r)); H = rng.uniform(-1, 1, size=(r, n)) / np.sqrt(n)
    bias = rng.uniform(-1, 1, size=(n,))
    layer = TLLayer(n, n, 0.5, initial_bias=bias, initial_lr_matrices=[G, H])
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32)
    y = run(layer, X, gy)
    np.savez_compressed(os.path.join(HERE, "tl_16.npz"), X=X, gy=gy, y=y, G=layer.G.detach().numpy(), H=layer.H.detach().numpy(),
                        bias=layer.bias.detach().numpy(), gG=layer.G.grad.numpy(), gH=layer.H.grad.numpy(), gbias=layer.bias.grad.numpy())
    print("golden fixtures written to", HERE)


def train_case():
    """Seeded data, model and numpy RNG of the training-loop fixture (shared with tests/test_training_helpers.py)."""
    rng = np.random.default_rng(77)
    X = rng.uniform(-1, 1, size=(230, 12)).astype(np.float32)
    Wt = rng.uniform(-1, 1, size=(12, 5)).astype(np.float32)
    y = (X @ Wt).argmax(1).astype(np.int64)
    Xv = rng.uniform(-1, 1, size=(40, 12)).astype(np.float32)
    yv = (Xv @ Wt).argmax(1).astype(np.int64)
    return X, y, Xv, yv


def train_fixture():
    """The reference's own training loop (training_helpers.train, :107-181) on a seeded torch.nn.Linear: the 9-tuple it returns and
    the trained weights, for two optimizers and both restore_best_model settings.  ``sklearn.utils.shuffle`` draws from numpy's
    global RNG, seeded here; the ragged last batch (230 = 4 x 50 + 30) is part of the case."""
    load_reference()
    from structurednets import training_helpers as RTH
    X, y, Xv, yv = train_case()
    out = {}
    for tag, opt, restore in (("sgd_best", torch.optim.SGD, True), ("adam_last", torch.optim.Adam, False)):
        torch.manual_seed(5)
        np.random.seed(11)
        model = torch.nn.Linear(12, 5)
        res = RTH.train(model, X, y, X_val=Xv, y_val=yv, patience=3, batch_size=50, lr=5e-2, restore_best_model=restore,
                        min_patience_improvement=1e-3, optimizer_class=opt)
        out[tag + "_start"] = np.asarray([float(v) for v in res[1:5]], dtype=np.float64)
        for name, h in zip(("tl", "ta", "vl", "va"), res[5:9]):
            out[tag + "_" + name] = np.asarray([float(v) for v in h], dtype=np.float64)
        out[tag + "_W"] = res[0].weight.detach().numpy().copy()
        out[tag + "_b"] = res[0].bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train_linear.npz"), **out)
    print("train fixture written:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_fixture()
    else:
        main()
        train_fixture()
