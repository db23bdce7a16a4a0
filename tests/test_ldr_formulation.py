"""CPU: the Krylov-stack formulation the LDR kernels implement (tests/ldr_emulator.py) equals the oracle -- weight matrix and
every gradient -- including the literal reference fixture (operators of norm > 1: all n powers) and a truncated series whose
length is not a multiple of the GEMM's 32-row boxes."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.ldr_layer import LDRLayer, LDR_REL_TOL
from tests import golden_io as GIO
from tests import ldr_emulator as E


def _run(rep, n, J, dW):
    A, B, G, H = rep
    bandA = E.band_of(A._indices().numpy(), A._values().detach().numpy(), n)
    bandB = E.band_of(B._indices().numpy(), B._values().detach().numpy(), n)
    bandBt = E.band_transpose(bandB, n)
    W, info = E.forward(bandA, bandBt, G.detach().numpy(), H.detach().numpy(), J, LDR_REL_TOL)
    gA, gB, gG, gH = E.backward(bandA, bandBt, info, dW, J)
    return W, info, gA, gB, gG, gH, bandA, bandB


def _band_grad_from_dense(g, n):
    band = np.zeros(3 * n + 2)
    idx = np.arange(n)
    band[n:2 * n] = g[idx, idx]
    band[1:n] = g[idx[1:], idx[:-1]]
    band[2 * n:3 * n - 1] = g[idx[:-1], idx[1:]]
    if n > 2:
        band[3 * n], band[3 * n + 1] = g[0, n - 1], g[n - 1, 0]
    return band


@pytest.mark.parametrize("n,share,J", [(12, 0.95, 12), (96, 0.3, 32), (40, 0.5, 32)])
def test_formulation_matches_oracle(n, share, J):
    if n == 12:
        z = GIO.load("ldr_12")
        rep = [t.detach().clone().requires_grad_(True) for t in GIO.ldr_rep(z)]
    else:
        np.random.seed(n)
        rep = [p.detach().clone().requires_grad_(True) for p in LDRLayer(n, n, share).representation_matrices]
    rng = np.random.default_rng(n)
    dW = rng.uniform(-1, 1, size=(n, n))
    Wo = O.ldr_weight(rep, (n, n))
    (Wo.double() * torch.tensor(dW)).sum().backward()
    W, info, gA, gB, gG, gH, _, _ = _run(rep, n, min(J, n), dW)
    assert info["conv"]
    if n != 12:
        assert info["terms"] < J and info["k_eff"] % 32 == 0
    scale = np.abs(Wo.detach().numpy()).max()
    assert np.abs(W - Wo.detach().numpy()).max() / scale < 1e-6
    dense = lambda t: (t.grad.to_dense() if t.grad.is_sparse else t.grad).numpy()
    for got, ref in ((gA, _band_grad_from_dense(dense(rep[0]), n)), (gB, _band_grad_from_dense(dense(rep[1]), n)),
                     (gG, rep[2].grad.numpy()), (gH, rep[3].grad.numpy())):
        assert np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30) < 1e-6
