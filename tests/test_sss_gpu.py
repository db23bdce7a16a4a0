"""Parity of the CUDA SSS path (through the C ABI) against the oracle -- needs a B200."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system
from tests.test_sss_plan import CASES, make, oracle_lists

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: 1e-5 relative in fp32 (relative to the largest entry of the compared tensor)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def run_case(layer, X, gy, chunk_note=""):
    dev = torch.device("cuda")
    lists = oracle_lists(layer)
    b = layer.bias.detach().clone().requires_grad_(True) if layer.use_bias else None
    yo = O.sss_forward(torch.tensor(X), *lists, b, layer.dims_in, layer.dims_out)
    (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(dev)
    y = layer(torch.tensor(X, device=dev))
    assert y.shape == yo.shape and y.is_contiguous()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    (y * torch.tensor(gy, device=dev)).sum().backward()
    worst = 0.0
    for li, name in enumerate("ABCDEFG"):
        got_all = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for p in getattr(layer, name)])
        ref_all = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)).reshape(-1) for p in lists[li]])
        if ref_all.size:
            e = rel_err(got_all, ref_all)
            worst = max(worst, e)
            assert e < RTOL, f"grad {name}: rel err {e:.3e} {chunk_note}"
    if b is not None:
        assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL
    return worst


@pytest.mark.parametrize("case", CASES)
def test_small_shapes(case, built_lib):
    layer, _, X = make(case)
    gy = np.random.default_rng(3).uniform(-1, 1, size=(case["B"], case["o"])).astype(np.float32)
    run_case(layer, X, gy)


@pytest.mark.parametrize("B", [256, 77])
def test_alexnet_last_layer_shape(B, built_lib):
    """BASELINE config C1: 4096 -> 1000, 500 stages, statespace dim 16 (B=77: ragged tile)."""
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=1001)
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm)
    assert layer.statespace_dim == 16
    rng = np.random.default_rng(1001)
    X = rng.uniform(-1, 1, size=(B, 4096)).astype(np.float32)
    gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
    run_case(layer, X, gy)


def test_no_bias_and_grad_accumulation(built_lib):
    layer, _, X = make(CASES[4], use_bias=False)
    dev = torch.device("cuda")
    layer = layer.to(dev)
    Xd = torch.tensor(X, device=dev)
    layer(Xd).sum().backward()
    g1 = layer.flat_grad().clone()
    layer(Xd).sum().backward()                     # accumulates like autograd
    torch.testing.assert_close(layer.flat_grad(), 2 * g1, rtol=1e-5, atol=1e-6)
    for p in layer.parameters():
        p.grad = None                              # optimizer.zero_grad(set_to_none=True)
    layer(Xd).sum().backward()
    torch.testing.assert_close(layer.flat_grad(), g1, rtol=1e-5, atol=1e-6)
    assert layer.A[1].grad.data_ptr() >= layer.flat_grad().data_ptr()


def test_linearity_at_full_batch(built_lib):
    """Size-independent property at a batch the oracle cannot do in seconds: the layer is affine in x."""
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=7)
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm).to("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    B = 8192
    x1 = torch.rand(B, 4096, device="cuda", generator=g) * 2 - 1
    x2 = torch.rand(B, 4096, device="cuda", generator=g) * 2 - 1
    with torch.no_grad():
        y1, y2, y12, y0 = layer(x1), layer(x2), layer(0.5 * x1 + 0.25 * x2), layer(torch.zeros_like(x1))
    lhs = y12 - y0
    rhs = 0.5 * (y1 - y0) + 0.25 * (y2 - y0)
    assert float((lhs - rhs).abs().max()) < 1e-5 * float(rhs.abs().max()) * 10
    torch.testing.assert_close(y0, layer.bias.detach().expand_as(y0))


def test_inplace_parameter_updates_are_seen_by_the_next_forward(built_lib):
    """Optimizers update parameters in place; the packed per-stage copy must follow (regression test)."""
    layer, _, X = make(CASES[4])
    layer = layer.to("cuda")
    Xd = torch.tensor(X, device="cuda")
    y0 = layer(Xd).detach().clone()
    with torch.no_grad():
        for p in layer.D:
            p.mul_(3.0)
        layer.B[2].add_(0.5)
    y1 = layer(Xd).detach()
    lists = oracle_lists(layer)
    yo = O.sss_forward(torch.tensor(X), *[[p.cpu() for p in l] for l in lists], layer.bias.detach().cpu(), layer.dims_in, layer.dims_out)
    assert float((y1 - y0).abs().max()) > 1e-2
    assert rel_err(y1.cpu().numpy(), yo.detach().numpy()) < RTOL
    opt = torch.optim.SGD(layer.parameters(), lr=0.05)
    before = float((layer(Xd) - 1).square().mean())
    for _ in range(20):
        opt.zero_grad(); loss = (layer(Xd) - 1).square().mean(); loss.backward(); opt.step()
    assert float((layer(Xd) - 1).square().mean()) < 0.95 * before
