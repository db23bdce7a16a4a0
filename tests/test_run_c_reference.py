"""The reference's own C implementation of the SSS forward (speed_comparison/run.c, compiled unmodified into oracle/_ref/run) as a
second checker: it must accept the oracle's output (CPU) and the CUDA kernels' output through the C ABI (gpu), and reject a
perturbed one.  run.c's own tolerance: 1e-3 absolute (run.c:177,357)."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from oracle import run_c
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system

CASES = [dict(i=76, o=14, n=7, d=5, ragged=True), dict(i=128, o=24, n=12, d=16, ragged=False), dict(i=320, o=64, n=40, d=16, ragged=False)]


def _binary_or_skip():
    if run_c.build_ref() is None:
        pytest.skip("oracle/_ref/run is not built (needs the reference source in the build container)")


def _layer(case, seed=3):
    sysm = random_mixed_system(case["i"], case["o"], case["n"], case["d"], seed=seed, ragged_state_dims=case["ragged"])
    rng = np.random.default_rng(seed)
    layer = SSSLayer(case["i"], case["o"], 0.9, initial_bias=rng.uniform(-1, 1, size=case["o"]), nb_states=case["n"], initial_system_approx=sysm)
    u = rng.uniform(-1, 1, size=(1, case["i"])).astype(np.float32)
    return layer, u


def _lists(layer):
    return [[p.detach().cpu().numpy() for p in getattr(layer, n)] for n in "ABCDEFG"]


@pytest.mark.parametrize("case", CASES)
def test_reference_c_program_accepts_the_oracle_forward(case):
    _binary_or_skip()
    layer, u = _layer(case)
    lists = [[p.detach() for p in getattr(layer, n)] for n in "ABCDEFG"]
    y = O.sss_forward(torch.tensor(u), *lists, layer.bias.detach(), layer.dims_in, layer.dims_out).numpy().reshape(-1)
    ok, out = run_c.run_reference_check(*_lists(layer), layer.bias.detach().numpy(), u, y)
    assert ok, out
    bad = y.copy()
    bad[len(bad) // 2] += 1e-2     # ten times run.c's tolerance: must be rejected
    ok2, out2 = run_c.run_reference_check(*_lists(layer), layer.bias.detach().numpy(), u, bad)
    assert not ok2 and "Checksum mismatch for the SSS layer" in out2


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["tc", "simt"])
@pytest.mark.parametrize("case", CASES[1:])
def test_reference_c_program_accepts_the_cuda_forward(case, path, built_lib, monkeypatch):
    _binary_or_skip()
    monkeypatch.setenv("SNB200_SSS_PATH", path)
    layer, u = _layer(case)
    params = _lists(layer)
    bias = layer.bias.detach().numpy().copy()
    layer = layer.to("cuda")
    rng = np.random.default_rng(9)
    X = rng.uniform(-1, 1, size=(40, case["i"])).astype(np.float32)
    X[17] = u[0]                                           # the checked vector sits inside a batch
    y = layer(torch.tensor(X, device="cuda")).detach().cpu().numpy()[17]
    ok, out = run_c.run_reference_check(*params, bias, u, y)
    assert ok, out
