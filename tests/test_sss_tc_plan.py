"""Chunked (tensor-core) formulation of the SSS layer, host side: the torch emulator of csrc/sss_tc.cu against the
oracle, and the chunk table SSSLayer hands to the C ABI (no GPU)."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system
from tests import sss_tc_emulator as E

TC_CASES = [
    dict(i=128, o=24, n=12, d=16, B=33, ragged=False),   # one chunk
    dict(i=64, o=16, n=4, d=8, B=20, ragged=True),       # ragged state dimensions
    dict(i=320, o=64, n=40, d=16, B=150, ragged=False),  # several chunks, chunk rows not 16-byte aligned
    dict(i=400, o=100, n=50, d=12, B=64, ragged=True),
]


def make_tc(case, seed=0, use_bias=True):
    sysm = random_mixed_system(case["i"], case["o"], case["n"], case["d"], seed=seed, ragged_state_dims=case["ragged"])
    rng = np.random.default_rng(seed + 1)
    bias = rng.uniform(-1, 1, size=(case["o"],)) if use_bias else None
    layer = SSSLayer(case["i"], case["o"], 0.9, use_bias=use_bias, initial_bias=bias, nb_states=case["n"], initial_system_approx=sysm)
    X = rng.uniform(-1, 1, size=(case["B"], case["i"])).astype(np.float32)
    return layer, X


def lists64(layer):
    return [[p.detach().double().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]


@pytest.mark.parametrize("case", TC_CASES)
def test_chunked_formulation_matches_oracle(case):
    layer, X = make_tc(case)
    l0, l1 = lists64(layer), lists64(layer)
    U = torch.tensor(X).double()
    b = layer.bias.detach().double()
    y0 = O.sss_forward(U, *l0, b, layer.dims_in, layer.dims_out)
    y1 = E.forward_chunked(U, *l1, b, layer.dims_in, layer.dims_out)
    assert float((y0 - y1).detach().abs().max()) < 1e-12
    (y0 ** 2).sum().backward()
    (y1 ** 2).sum().backward()
    for a, c in zip(sum(l0, []), sum(l1, [])):
        if a.grad is not None and a.numel():
            assert float((a.grad - c.grad).abs().max()) < 1e-10


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_chunk_table(case):
    layer, _ = make_tc(case)
    host = layer.build_tc_host_plan()
    assert host is not None
    ch = host["chunks"]
    assert [(int(a), int(b)) for a, b in ch[:, :2]] == E.make_chunks(layer.dims_in, layer.dims_out)
    assert ch[0, 0] == 0 and ch[-1, 1] == layer.nb_states and np.all(ch[1:, 0] == ch[:-1, 1])
    assert np.all(ch[:, 3] <= 160) and np.all(ch[:, 5] <= 32) and np.all(ch[:, 1] - ch[:, 0] <= 16)
    assert int(ch[:, 3].sum()) == layer.input_dim and int(ch[:, 5].sum()) == layer.output_dim
    assert np.all(ch[:, 6] == (ch[:, 3] + 31) // 32)


def test_layers_outside_the_tc_limits_use_the_simt_path():
    sysm = random_mixed_system(76, 14, 10, 3, seed=0)
    layer = SSSLayer(76, 14, 0.9, nb_states=10, initial_system_approx=sysm)   # 14 outputs: not a multiple of 4
    assert layer.build_tc_host_plan() is None and not layer._use_tc_path()
    sysm = random_mixed_system(128, 32, 4, 20, seed=0)
    layer = SSSLayer(128, 32, 0.9, nb_states=4, initial_system_approx=sysm)   # state dimension 20 > 16
    assert layer.build_tc_host_plan() is None
