"""Hankel-SVD system identification (structurednets_b200/sss_identification.py): the from-dense constructor path of SSSLayer without
tvsclib.  Parity with tvsclib is unpinned by the reference (realisations are unique only up to a
similarity); pinned here: exact reconstruction without truncation, exact recovery of a semiseparable matrix at its order, monotone
approximation error, and the reference's own known answer "layer forward == x to_matrix()^T + b" (tests/test_layers.py:118-121)."""
import numpy as np
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.sss_layer import SSSLayer, get_nb_parameters
from structurednets_b200.sss_identification import identify_mixed_system
from structurednets_b200.synth import random_mixed_system, standard_dims


def test_exact_without_truncation_and_monotone_with_it():
    rng = np.random.default_rng(0)
    T = rng.uniform(-1, 1, size=(30, 48))
    di, do = standard_dims(48, 30, 6)
    full = identify_mixed_system(T, di, do, 100)
    assert np.abs(full.to_matrix() - T).max() < 1e-12
    errs = [np.linalg.norm(identify_mixed_system(T, di, do, d).to_matrix() - T) for d in range(0, 14)]
    assert all(b <= a + 1e-12 for a, b in zip(errs[:-1], errs[1:])) and errs[-1] < errs[0]
    # d = 0: only the block diagonal survives
    diag = identify_mixed_system(T, di, do, 0).to_matrix()
    io, oo = np.concatenate([[0], np.cumsum(di)]), np.concatenate([[0], np.cumsum(do)])
    for k in range(6):
        np.testing.assert_array_equal(diag[oo[k]:oo[k + 1], io[k]:io[k + 1]], T[oo[k]:oo[k + 1], io[k]:io[k + 1]])
    assert np.count_nonzero(diag) == sum(int(a) * int(b) for a, b in zip(di, do))


def test_recovers_a_semiseparable_matrix_at_its_order():
    sysm = random_mixed_system(76, 14, 7, 5, seed=3, ragged_state_dims=True)
    T = sysm.to_matrix()
    rec = identify_mixed_system(T, sysm.dims_in, sysm.dims_out, 5)
    assert np.abs(rec.to_matrix() - T).max() < 1e-12
    for a, b in zip(rec.causal_system.stages, sysm.causal_system.stages):
        assert a.A_matrix.shape[0] <= 5 and a.B_matrix.shape[1] == b.B_matrix.shape[1] and a.D_matrix.shape == b.D_matrix.shape


def test_layer_from_dense_matrix_without_tvsclib():
    """reference test_sss_layer (tests/test_layers.py:106-121): the layer built from a dense matrix reproduces its own
    initial_weight_matrix; checked through the CPU oracle (the CUDA forward is checked against the same oracle in the gpu tests)."""
    rng = np.random.default_rng(1)
    W = rng.uniform(-1, 1, size=(14, 76))
    layer = SSSLayer(76, 14, 0.6, initial_weight_matrix=W, nb_states=7)
    assert layer.state_matrices_initialized and layer.initial_weight_matrix.shape == (14, 76)
    assert sum(p.numel() for n, p in layer.named_parameters() if n != "bias") <= get_nb_parameters((14, 76), layer.statespace_dim, 7)
    assert list(layer.dims_in) == [11, 11, 11, 11, 11, 11, 10] and list(layer.dims_out) == [2] * 7
    for k in range(7):   # the shapes the reference documents: state dimensions bounded by statespace_dim
        assert layer.A[k].shape[0] <= layer.statespace_dim and layer.A[k].shape[1] <= layer.statespace_dim
        assert layer.B[k].shape == (layer.A[k].shape[0], layer.dims_in[k]) and layer.C[k].shape == (2, layer.A[k].shape[1])
        assert layer.F[k].shape == (layer.E[k].shape[0], layer.dims_in[k]) and layer.G[k].shape == (2, layer.E[k].shape[1])
    X = rng.uniform(-1, 1, size=(9, 76)).astype(np.float32)
    lists = [[p.detach() for p in getattr(layer, n)] for n in "ABCDEFG"]
    y = O.sss_forward(torch.tensor(X), *lists, layer.bias.detach(), layer.dims_in, layer.dims_out).numpy()
    ref = X @ layer.initial_weight_matrix.T.astype(np.float32) + layer.bias.detach().numpy()
    assert np.abs(y - ref).max() < 1e-5
    # more budget -> a better approximation of W (and the full budget reproduces it)
    err = lambda share: np.linalg.norm(SSSLayer(76, 14, share, initial_weight_matrix=W, nb_states=7).initial_weight_matrix - W)
    assert err(0.9) < err(0.6) < err(0.3)
