"""Parity of the CUDA LR / PSM / H-matrix / LDR / TL paths (through the C ABI) against the golden fixtures
of the unmodified reference and against the oracle -- needs a B200."""
import pickle

import numpy as np
import pytest
import scipy.sparse
import torch

from oracle import layers_cpu as O
from structurednets_b200.hmatrix import HMatrix, build_hmat_block_cluster_tree
from structurednets_b200.layers.hmat_layer import HMatLayer
from structurednets_b200.layers.ldr_layer import LDRLayer
from structurednets_b200.layers.lr_layer import LRLayer
from structurednets_b200.layers.psm_layer import PSMLayer
from structurednets_b200.layers.tl_layer import TLLayer
from tests import golden_io as GIO

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-5  # fp32 bar of north_star, relative to the largest entry of the compared tensor


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if b.size else 0.0


def dense_grad(p):
    g = p.grad
    return (g.to_dense() if g.is_sparse else g).detach().cpu().numpy()


def fwd_bwd(layer, z):
    layer = layer.to(DEV)
    y = layer(torch.tensor(z["X"], device=DEV))
    (y * torch.tensor(z["gy"], device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), z["y"]) < RTOL
    if layer.use_bias:
        assert rel_err(layer.bias.grad.cpu().numpy(), z["gbias"]) < RTOL
    return layer


def test_lr_matches_reference_fixture(built_lib):
    z = GIO.load("lr_96x40")
    layer = fwd_bwd(LRLayer(96, 40, 0.5, initial_bias=z["bias"], initial_lr_components=[z["left"], z["right"]]), z)
    assert rel_err(dense_grad(layer.left_lr), z["gleft"]) < RTOL
    assert rel_err(dense_grad(layer.right_lr), z["gright"]) < RTOL
    assert list(layer.state_dict().keys()) == ["bias", "left_lr", "right_lr"]


def test_lr_resnet_shape_fp32_vs_oracle(built_lib):
    rng = np.random.default_rng(2002)
    B, i, o, r = 512, 2048, 1000, 128
    layer = LRLayer(i, o, 0.1906)
    assert layer.left_lr.shape == (o, r) and layer.right_lr.shape == (r, i)   # BASELINE C2: rank 128
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32) / B
    L, R, b = [t.detach().clone().requires_grad_(True) for t in (layer.left_lr, layer.right_lr, layer.bias)]
    yo = O.lr_forward(torch.tensor(X), L, R, b); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    assert rel_err(dense_grad(layer.left_lr), L.grad.numpy()) < RTOL
    assert rel_err(dense_grad(layer.right_lr), R.grad.numpy()) < RTOL
    assert rel_err(dense_grad(layer.bias), b.grad.numpy()) < RTOL


def test_psm_two_factors_matches_reference_fixture(built_lib):
    z = GIO.load("psm_50x30_2f")
    f = [scipy.sparse.csr_matrix(z["f0_dense"]), scipy.sparse.csr_matrix(z["f1_dense"])]
    layer = fwd_bwd(PSMLayer(50, 30, initial_bias=z["bias"], sparse_matrices=f), z)
    for k, p in enumerate(layer.sparse_matrices):
        assert p.is_sparse and p.grad.is_sparse and p.grad._nnz() == p._nnz()      # sparse grads on the same pattern
        assert rel_err(dense_grad(p), z[f"gf{k}_dense"]) < RTOL
    assert list(layer.state_dict().keys()) == ["bias", "sparse_matrices.0", "sparse_matrices.1"]


@pytest.mark.parametrize("path", ["sparse", "dense"])
def test_psm_three_factors_intended_order_and_sgd_step(built_lib, monkeypatch, path):
    """3 factors, rectangular -- compared with the corrected-order restatement; both CUDA paths (sparse chain with
    the intermediates in shared memory / batch-independent dense product on the tensor cores)."""
    monkeypatch.setenv("SNB200_PSM_PATH", path)
    rng = np.random.default_rng(3003)
    i, o, B = 96, 40, 37
    mx = max(i, o)
    S = [scipy.sparse.random(o, mx, density=0.1, random_state=1, format="csr"), scipy.sparse.random(mx, mx, density=0.1, random_state=2, format="csr"),
         scipy.sparse.random(mx, i, density=0.1, random_state=3, format="csr")]
    layer = PSMLayer(i, o, sparse_matrices=S)
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32)
    dense = [torch.tensor(s.toarray()).float().requires_grad_(True) for s in S]
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.psm_forward(torch.tensor(X), dense, b); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    Xd, gyd = torch.tensor(X, device=DEV), torch.tensor(gy, device=DEV)
    y = layer(Xd); (y * gyd).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    for k, p in enumerate(layer.sparse_matrices):
        mask = S[k].toarray() != 0
        assert rel_err(dense_grad(p), dense[k].grad.numpy() * mask) < RTOL
    # the reference's training recipe for sparse parameters: plain SGD (psm_layer.py:14-15)
    opt = torch.optim.SGD(layer.parameters(), lr=0.1)
    before = float((layer(Xd) - 1).square().mean())
    for _ in range(5):
        opt.zero_grad(); loss = (layer(Xd) - 1).square().mean(); loss.backward(); opt.step()
    assert float((layer(Xd) - 1).square().mean()) < before
    assert all(p._nnz() == s.nnz for p, s in zip(layer.sparse_matrices, S))


def test_psm_dense_product_path_large_batch_vs_oracle(built_lib, monkeypatch):
    """The path a large batch takes by default (B >= 1024): W^T multiplied out once, tensor-core GEMMs, gradient projected onto
    every factor's pattern -- against the oracle, and against the sparse chain kernel on the same inputs."""
    rng = np.random.default_rng(3004)
    i, o, B, mx = 512, 200, 2048, 512
    S = [scipy.sparse.random(o, mx, density=0.05, random_state=4, format="csr"), scipy.sparse.random(mx, mx, density=0.05, random_state=5, format="csr"),
         scipy.sparse.random(mx, i, density=0.05, random_state=6, format="csr")]
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32) / B
    ref = PSMLayer(i, o, sparse_matrices=S)
    dense = [torch.tensor(s.toarray()).float().requires_grad_(True) for s in S]
    b = ref.bias.detach().clone().requires_grad_(True)
    yo = O.psm_forward(torch.tensor(X), dense, b); (yo * torch.tensor(gy)).sum().backward()
    outs = {}
    for path in ("dense", "sparse"):
        monkeypatch.setenv("SNB200_PSM_PATH", path)
        layer = PSMLayer(i, o, sparse_matrices=S, initial_bias=ref.bias.detach().numpy()).to(DEV)
        assert layer.use_dense_path(B) == (path == "dense")
        y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
        assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
        assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL
        for k, p in enumerate(layer.sparse_matrices):
            assert p.grad.is_sparse and p.grad._nnz() == p._nnz()
            assert rel_err(dense_grad(p), dense[k].grad.numpy() * (S[k].toarray() != 0)) < RTOL
        outs[path] = y.detach().cpu().numpy()
    monkeypatch.delenv("SNB200_PSM_PATH")
    assert PSMLayer(i, o, sparse_matrices=S).use_dense_path(B) and not PSMLayer(i, o, sparse_matrices=S).use_dense_path(64)
    assert rel_err(outs["dense"], outs["sparse"]) < RTOL


def _hmat_from_fixture(z):
    tree = build_hmat_block_cluster_tree((40, 64), eta=0.5, min_block_size=2)
    leaves = {(l.row_range.start, l.row_range.stop, l.col_range.start, l.col_range.stop): l for l in tree.get_all_leaf_elements()}
    for (r0, r1, c0, c1, L, R) in GIO.hmat_components(z):
        leaves[(r0, r1, c0, c1)].set_hmatrix_component(L, R)
    return HMatrix(tree, shape=(40, 64))


def test_hmat_matches_reference_fixture(built_lib):
    z = GIO.load("hmat_40x64")
    hm = _hmat_from_fixture(z)
    layer = HMatLayer(64, 40, 0.9, initial_bias=z["bias"], initial_hmatrix=hm)
    assert len(layer.hmatrix_components) == int(z["ncomp"])
    for c, comp in enumerate(layer.hmatrix_components):      # same leaf order as the reference's traversal
        assert [comp.row_range.start, comp.row_range.stop, comp.col_range.start, comp.col_range.stop] == [int(v) for v in z[f"rng{c}"]]
    layer = fwd_bwd(layer, z)
    for c, comp in enumerate(layer.hmatrix_components):
        assert rel_err(dense_grad(comp.left_lr), z[f"gL{c}"]) < 2 * RTOL
        assert rel_err(dense_grad(comp.right_lr), z[f"gR{c}"]) < 2 * RTOL
    keys = list(layer.state_dict().keys())
    assert keys[:3] == ["bias", "hmatrix_components.0.left_lr", "hmatrix_components.0.right_lr"]
    clone = pickle.loads(pickle.dumps(layer))                # training_helpers.py:132 snapshots models by pickle
    torch.testing.assert_close(clone(torch.tensor(z["X"], device=DEV)), layer(torch.tensor(z["X"], device=DEV)))


def test_hmat_inception_shape_vs_oracle(built_lib):
    """BASELINE C4-H: 2048 -> 1000, eta 0.5, min block 2 (2560 leaves), rank min(6, dim-1)."""
    rng = np.random.default_rng(4004)
    tree = build_hmat_block_cluster_tree((1000, 2048), eta=0.5, min_block_size=2)
    for leaf in tree.get_all_leaf_elements():
        rows, cols = len(leaf.row_range), len(leaf.col_range)
        k = min(6, min(rows, cols) - 1)
        leaf.set_hmatrix_component(torch.tensor(rng.uniform(-1, 1, (rows, k)) / np.sqrt(k)).float(), torch.tensor(rng.uniform(-1, 1, (k, cols)) / np.sqrt(cols)).float())
    layer = HMatLayer(2048, 1000, 0.2, initial_hmatrix=HMatrix(tree, shape=(1000, 2048)))
    assert len(layer.hmatrix_components) == 2560
    B = 48
    X = rng.uniform(-1, 1, size=(B, 2048)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
    comps = [(c.row_range.start, c.row_range.stop, c.col_range.start, c.col_range.stop, c.left_lr.detach().clone().requires_grad_(True),
              c.right_lr.detach().clone().requires_grad_(True)) for c in layer.hmatrix_components]
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.hmat_forward(torch.tensor(X), comps, b, 1000); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    gl = np.concatenate([dense_grad(c.left_lr).ravel() for c in layer.hmatrix_components]); gl_o = np.concatenate([c[4].grad.numpy().ravel() for c in comps])
    gr = np.concatenate([dense_grad(c.right_lr).ravel() for c in layer.hmatrix_components]); gr_o = np.concatenate([c[5].grad.numpy().ravel() for c in comps])
    assert rel_err(gl, gl_o) < RTOL and rel_err(gr, gr_o) < RTOL


def test_ldr_matches_reference_fixture(built_lib):
    """The fixture comes from the reference's literal matrix_power construction with operators of norm > 1,
    so all n Krylov terms matter.  LDR tolerance (stated separately): 1e-4 relative -- the
    reference accumulates float64 terms into a float32 matrix, we accumulate in float64 and round once."""
    z = GIO.load("ldr_12")
    rep = GIO.ldr_rep(z)
    layer = LDRLayer(12, 12, 0.95, initial_bias=z["bias"], initial_representation_matrices=rep).to(DEV)
    y = layer(torch.tensor(z["X"], device=DEV)); (y * torch.tensor(z["gy"], device=DEV)).sum().backward()
    assert layer.last_nb_terms == 12
    tol = 1e-4
    assert rel_err(y.detach().cpu().numpy(), z["y"]) < tol
    A, Bm, G, H = layer.representation_matrices
    assert A.dtype == torch.float64 and A.is_sparse and A.grad.is_sparse and G.grad.dtype == torch.float64
    assert rel_err(dense_grad(A), z["gA_dense"]) < tol and rel_err(dense_grad(Bm), z["gB_dense"]) < tol
    assert rel_err(dense_grad(G), z["gG"]) < tol and rel_err(dense_grad(H), z["gH"]) < tol
    assert rel_err(layer.bias.grad.cpu().numpy(), z["gbias"]) < tol
    assert list(layer.state_dict().keys()) == ["bias"] + [f"representation_matrices.{i}" for i in range(4)]


def test_ldr_converging_series_vs_oracle(built_lib):
    """Glorot-scale operators (the reference's own init): the series converges long before n terms."""
    np.random.seed(5)
    n, B = 96, 21
    layer = LDRLayer(n, n, 0.3)
    rng = np.random.default_rng(5005)
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32)
    rep = [p.detach().clone().requires_grad_(True) for p in layer.representation_matrices]
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.ldr_forward(torch.tensor(X), rep, b, (n, n)); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert layer.last_nb_terms < n
    tol = 1e-4
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < tol
    A, Bm, G, H = layer.representation_matrices
    mA = rep[0].detach().to_dense().numpy() != 0; mB = rep[1].detach().to_dense().numpy() != 0
    gA_o = (rep[0].grad.to_dense() if rep[0].grad.is_sparse else rep[0].grad).numpy() * mA
    gB_o = (rep[1].grad.to_dense() if rep[1].grad.is_sparse else rep[1].grad).numpy() * mB
    assert rel_err(dense_grad(A), gA_o) < tol and rel_err(dense_grad(Bm), gB_o) < tol
    assert rel_err(dense_grad(G), rep[2].grad.numpy()) < tol and rel_err(dense_grad(H), rep[3].grad.numpy()) < tol


def test_tl_matches_reference_fixture(built_lib):
    z = GIO.load("tl_16")
    layer = fwd_bwd(TLLayer(16, 16, 0.5, initial_bias=z["bias"], initial_lr_matrices=[z["G"], z["H"]]), z)
    assert rel_err(dense_grad(layer.G), z["gG"]) < RTOL and rel_err(dense_grad(layer.H), z["gH"]) < RTOL
    assert list(layer.state_dict().keys()) == ["bias", "G", "H"]


def test_tl_larger_vs_oracle(built_lib):
    rng = np.random.default_rng(6006)
    n, r, B = 200, 5, 33
    G = rng.uniform(-1, 1, (n, r)) / np.sqrt(n); H = rng.uniform(-1, 1, (r, n)) / np.sqrt(n)
    layer = TLLayer(n, n, 0.5, initial_lr_matrices=[G, H])
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32)
    Gt, Ht, b = [t.detach().clone().requires_grad_(True) for t in (layer.G, layer.H, layer.bias)]
    yo = O.tl_forward(torch.tensor(X), Gt, Ht, b); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    assert rel_err(dense_grad(layer.G), Gt.grad.numpy()) < 2 * RTOL and rel_err(dense_grad(layer.H), Ht.grad.numpy()) < 2 * RTOL


def test_degenerate_layers_match_reference_conventions(built_lib):
    ldr = LDRLayer(4, 4, 0.1)                 # displacement rank < 0 -> dummy parameter, zero output (ldr_layer.py:25-27,51-52)
    assert ldr.representation_matrices is None and float(ldr(torch.ones(2, 4)).abs().sum()) == 0.0
    tl = TLLayer(6, 6, 0.01).to(DEV)          # rank 0 -> bias only (tl_layer.py:62-66)
    out = tl(torch.ones(3, 6, device=DEV))
    torch.testing.assert_close(out, tl.bias.detach().expand(3, 6))
    with pytest.raises(AssertionError):
        LDRLayer(6, 5, 0.5)                   # square only
