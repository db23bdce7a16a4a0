"""Host logic of the SSS layer (no GPU): constructor contract, plan tables, pickling."""
import pickle

import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.sss_layer import SSSLayer, get_nb_parameters, standard_dims_in_dims_out_computation
from structurednets_b200.synth import random_mixed_system
from tests import plan_emulator as E

CASES = [
    dict(i=76, o=14, n=10, d=3, B=51, ragged=False),    # reference tests/test_layers.py:106-121 shape
    dict(i=50, o=50, n=5, d=4, B=10, ragged=True),      # tests/test_layers.py:201-207 shape
    dict(i=31, o=20, n=7, d=5, B=9, ragged=True),       # odd stage count
    dict(i=12, o=9, n=2, d=2, B=5, ragged=False),       # two stages: every stage is a boundary stage
    dict(i=128, o=24, n=12, d=16, B=33, ragged=False),  # 9+ wide / 2-out stages like the 4096->1000 config
]


def make(case, seed=0, use_bias=True):
    sysm = random_mixed_system(case["i"], case["o"], case["n"], case["d"], seed=seed, ragged_state_dims=case["ragged"])
    rng = np.random.default_rng(seed + 1)
    bias = rng.uniform(-1, 1, size=(case["o"],)) if use_bias else None
    layer = SSSLayer(case["i"], case["o"], 0.9, use_bias=use_bias, initial_bias=bias, nb_states=case["n"], initial_system_approx=sysm)
    X = rng.uniform(-1, 1, size=(case["B"], case["i"])).astype(np.float32)
    return layer, sysm, X


def oracle_lists(layer):
    return [[p.detach().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]


@pytest.mark.parametrize("case", CASES)
def test_plan_forward_backward_matches_oracle(case):
    layer, sysm, X = make(case)
    stages, chunks, meta = layer.build_host_plan(chunk_len=3)
    flat = layer.flat_parameters().detach().numpy().copy()
    bias = layer.bias.detach().numpy()
    y, ckpt = E.forward(stages, chunks, meta, flat, X, bias)
    lists = oracle_lists(layer)
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.sss_forward(torch.tensor(X), *lists, b, layer.dims_in, layer.dims_out)
    np.testing.assert_allclose(y, yo.detach().numpy(), rtol=1e-5, atol=1e-5)
    # dense-matrix cross-check (what the reference's own test_sss_layer pins)
    T = sysm.to_matrix()
    np.testing.assert_allclose(y, X @ T.T.astype(np.float32) + bias, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(O.sss_to_matrix(*[[p.detach() for p in l] for l in lists], layer.dims_in, layer.dims_out), T, atol=1e-6)
    # backward
    rng = np.random.default_rng(5)
    gy = rng.uniform(-1, 1, size=y.shape).astype(np.float32)
    (yo * torch.tensor(gy)).sum().backward()
    g, gb = E.backward(stages, chunks, meta, flat, X, gy, ckpt)
    offs = layer._param_offsets()
    for li, name in enumerate("ABCDEFG"):
        for k, p in enumerate(lists[li]):
            o = offs[(name, k)]
            ref = p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)  # unused params get no grad
            np.testing.assert_allclose(g[o:o + p.numel()].reshape(p.shape), ref, rtol=1e-4, atol=1e-4, err_msg=f"{name}[{k}]")
    np.testing.assert_allclose(gb, b.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_state_dict_keys_and_shapes_match_reference_layout():
    layer, _, _ = make(CASES[0])
    keys = list(layer.state_dict().keys())
    assert keys[0] == "bias"
    n = CASES[0]["n"]
    assert keys[1:] == [f"{name}.{k}" for name in "ABCDEFG" for k in range(n)]
    # boundary shapes of SURVEY.md section 3b
    assert tuple(layer.A[0].shape) == (3, 0) and tuple(layer.E[0].shape) == (0, 3)
    assert tuple(layer.A[n - 1].shape) == (0, 3) and tuple(layer.G[n - 1].shape)[1] == 0
    # every parameter is a view into one flat buffer, in state_dict order
    flat = layer.flat_parameters()
    off = 0
    for k in keys:
        p = dict(layer.named_parameters())[k]
        if p.numel():
            assert p.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    assert off == flat.numel()


def test_pickle_roundtrip_is_compact_and_keeps_views():
    layer, _, _ = make(CASES[1])
    blob = pickle.dumps(layer)
    assert len(blob) < 200_000
    clone = pickle.loads(blob)
    for (k1, p1), (k2, p2) in zip(layer.named_parameters(), clone.named_parameters()):
        assert k1 == k2 and torch.equal(p1, p2)
    assert clone._flat_is_valid()
    with torch.no_grad():
        clone.A[1].add_(1.0)
    assert not torch.equal(clone.A[1], layer.A[1])      # independent storage
    assert clone.flat_parameters()._version > 0           # views share the flat buffer's version counter


def test_helper_functions_match_reference_formulas():
    di, do = standard_dims_in_dims_out_computation(4096, 1000, 500)
    assert list(di[:96]) == [9] * 96 and list(di[96:]) == [8] * 404 and list(do) == [2] * 500
    assert get_nb_parameters((1000, 4096), 16, 500) == 426240


def test_cpu_input_raises_no_fallback():
    layer, _, X = make(CASES[0])
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.tensor(X))


def test_budget_without_room_returns_zeros_like_reference():
    layer = SSSLayer(20, 16, 0.01, nb_states=4)   # statespace_dim 0 and not even D fits: sss_layer.py:130-131
    assert not getattr(layer, "state_matrices_initialized", False)
    out = layer(torch.zeros(3, 20))
    assert out.shape == (3, 16) and float(out.abs().sum()) == 0.0
    assert layer.get_nb_parameters() == 0
