"""Pins the oracle (oracle/layers_cpu.py) against the golden fixtures produced by the unmodified
reference modules (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from tests import golden_io as GIO

TOL = dict(rtol=1e-5, atol=1e-5)


def leaf(t):
    return t.detach().clone().requires_grad_(True)


@pytest.mark.parametrize("name", ["sss_76x14", "sss_50x50"])
def test_sss_oracle_matches_reference(name):
    z = GIO.load(name)
    lists = [[leaf(p) for p in l] for l in GIO.sss_lists(z)]
    b = leaf(torch.tensor(z["bias"]))
    y = O.sss_forward(torch.tensor(z["X"]), *lists, b, z["dims_in"], z["dims_out"])
    np.testing.assert_allclose(y.detach().numpy(), z["y"], **TOL)
    (y * torch.tensor(z["gy"])).sum().backward()
    for l, gl in zip(lists, GIO.sss_grad_lists(z)):
        for p, g in zip(l, gl):
            got = p.grad.numpy() if p.grad is not None else np.zeros_like(g)
            np.testing.assert_allclose(got, g, **TOL)
    np.testing.assert_allclose(b.grad.numpy(), z["gbias"], **TOL)
    # dense block formula == what the layer computes (the reference's own test_sss_layer known answer)
    T = O.sss_to_matrix(*GIO.sss_lists(z), z["dims_in"], z["dims_out"])
    np.testing.assert_allclose(z["X"] @ T.T + z["bias"], z["y"], rtol=1e-4, atol=1e-4)


def test_lr_oracle_matches_reference():
    z = GIO.load("lr_96x40")
    L, R, b = leaf(torch.tensor(z["left"])), leaf(torch.tensor(z["right"])), leaf(torch.tensor(z["bias"]))
    y = O.lr_forward(torch.tensor(z["X"]), L, R, b)
    np.testing.assert_allclose(y.detach().numpy(), z["y"], **TOL)
    (y * torch.tensor(z["gy"])).sum().backward()
    np.testing.assert_allclose(L.grad.numpy(), z["gleft"], **TOL)
    np.testing.assert_allclose(R.grad.numpy(), z["gright"], **TOL)
    np.testing.assert_allclose(b.grad.numpy(), z["gbias"], **TOL)


def test_psm_oracle_matches_reference_two_factors():
    z = GIO.load("psm_50x30_2f")
    f = [leaf(torch.tensor(z["f0_dense"]).float()), leaf(torch.tensor(z["f1_dense"]).float())]
    b = leaf(torch.tensor(z["bias"]))
    for fwd in (O.psm_forward, O.psm_forward_literal):   # identical for 2 factors
        for t in f + [b]:
            t.grad = None
        y = fwd(torch.tensor(z["X"]), f, b)
        np.testing.assert_allclose(y.detach().numpy(), z["y"], **TOL)
        (y * torch.tensor(z["gy"])).sum().backward()
        for k in range(2):
            mask = z[f"f{k}_dense"] != 0      # the reference's sparse grads live on the COO pattern only
            np.testing.assert_allclose(f[k].grad.numpy() * mask, z[f"gf{k}_dense"], **TOL)
        np.testing.assert_allclose(b.grad.numpy(), z["gbias"], **TOL)


def test_psm_intended_order_three_factors_is_the_product():
    rng = np.random.default_rng(3)
    S = [rng.uniform(-1, 1, size=s) * (rng.uniform(size=s) < 0.3) for s in ((7, 12), (12, 12), (12, 9))]
    X = rng.uniform(-1, 1, size=(5, 9))
    y = O.psm_forward(torch.tensor(X), [torch.tensor(s) for s in S], None)
    np.testing.assert_allclose(y.numpy(), X @ (S[0] @ S[1] @ S[2]).T, rtol=1e-10, atol=1e-12)   # psm_approximator.py:96-105


def test_hmat_oracle_matches_reference():
    z = GIO.load("hmat_40x64")
    comps = [(r0, r1, c0, c1, leaf(L), leaf(R)) for (r0, r1, c0, c1, L, R) in GIO.hmat_components(z)]
    b = leaf(torch.tensor(z["bias"]))
    y = O.hmat_forward(torch.tensor(z["X"]), comps, b, 40)
    np.testing.assert_allclose(y.detach().numpy(), z["y"], **TOL)
    (y * torch.tensor(z["gy"])).sum().backward()
    for c, comp in enumerate(comps):
        np.testing.assert_allclose(comp[4].grad.numpy(), z[f"gL{c}"], **TOL)
        np.testing.assert_allclose(comp[5].grad.numpy(), z[f"gR{c}"], **TOL)
    np.testing.assert_allclose(z["X"] @ O.hmat_to_dense(comps, (40, 64)).T + z["bias"], z["y"], rtol=1e-4, atol=1e-4)


def test_ldr_oracle_matches_reference_literal_and_recurrence():
    z = GIO.load("ldr_12")
    for literal in (True, False):
        rep = [leaf(t) for t in GIO.ldr_rep(z)]
        b = leaf(torch.tensor(z["bias"]))
        y = O.ldr_forward(torch.tensor(z["X"]), rep, b, (12, 12), literal=literal)
        np.testing.assert_allclose(y.detach().numpy(), z["y"], **TOL)
        (y * torch.tensor(z["gy"])).sum().backward()
        gA = rep[0].grad.to_dense().numpy() if rep[0].grad.is_sparse else rep[0].grad.numpy()
        gB = rep[1].grad.to_dense().numpy() if rep[1].grad.is_sparse else rep[1].grad.numpy()
        maskA = GIO.ldr_rep(z)[0].to_dense().numpy() != 0
        np.testing.assert_allclose(gA * maskA, z["gA_dense"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(gB * maskA.T * 0 + gB * (GIO.ldr_rep(z)[1].to_dense().numpy() != 0), z["gB_dense"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(rep[2].grad.numpy(), z["gG"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(rep[3].grad.numpy(), z["gH"], rtol=1e-4, atol=1e-5)


def test_tl_oracle_matches_reference():
    z = GIO.load("tl_16")
    G, H, b = leaf(torch.tensor(z["G"])), leaf(torch.tensor(z["H"])), leaf(torch.tensor(z["bias"]))
    y = O.tl_forward(torch.tensor(z["X"]), G, H, b)
    np.testing.assert_allclose(y.detach().numpy(), z["y"], **TOL)
    (y * torch.tensor(z["gy"])).sum().backward()
    np.testing.assert_allclose(G.grad.numpy(), z["gG"], **TOL)
    np.testing.assert_allclose(H.grad.numpy(), z["gH"], **TOL)
