"""Parity at (or near) the sizes BASELINE.json's configs name, through the module API on a B200 -- the cases the small
fixtures cannot see: PSM C3 (4096 -> 1000, three factors, 409 599 non-zeros) on both CUDA paths, the SSS chain kernels
(tcgen05 chunk scans) at C1's shape and batch, the H-matrix leaf-by-leaf kernels, LDR at n = 512 / 2048 and a Toeplitz-like
layer whose Krylov stack has more than 65 535 rows.  Tolerances: 1e-5 relative to the largest entry of the compared tensor
for the fp32 layers (north_star), 1e-4 for LDR (stated in DESIGN.md)."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200.hmatrix import HMatrix, build_hmat_block_cluster_tree
from structurednets_b200.layers.hmat_layer import HMatLayer
from structurednets_b200.layers.ldr_layer import LDRLayer
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.layers.tl_layer import TLLayer
from structurednets_b200.synth import random_mixed_system

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-5


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if b.size else 0.0


def rms_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30)) if b.size else 0.0


def dense_grad(p):
    g = p.grad
    return (g.to_dense() if g.is_sparse else g).detach().cpu().numpy()


@pytest.mark.parametrize("path", ["sparse", "dense"])
def test_psm_c3_size_vs_oracle(built_lib, monkeypatch, path):
    """BASELINE C3's layer exactly as bench.py builds it (4096 -> 1000, S0 1000x4096, S1, S2 4096x4096, 136 533 non-zeros each),
    batch 1024: the oracle's dense chain (psm_layer.py:51-58 in the intended order) finishes in seconds."""
    from bench import PSMWorkload
    monkeypatch.setenv("SNB200_PSM_PATH", path)
    wl = PSMWorkload()
    S = wl.factors()
    assert sum(s.nnz for s in S) == 409599
    B = 1024
    rng = np.random.default_rng(3333)
    X = rng.uniform(-1, 1, size=(B, 4096)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
    layer = wl.make_layer("cpu")
    dense = [torch.tensor(s.toarray()).float().requires_grad_(True) for s in S]
    b = layer.bias.detach().clone().requires_grad_(True)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    yo = O.psm_forward(torch.tensor(X), dense, b); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    assert layer.use_dense_path(B) == (path == "dense")
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    assert rms_rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL
    for k, p in enumerate(layer.sparse_matrices):
        assert p.grad.is_sparse and p.grad._nnz() == p._nnz()
        ref = dense[k].grad.numpy() * (S[k].toarray() != 0)
        assert rel_err(dense_grad(p), ref) < RTOL, "factor %d" % k
        assert rms_rel_err(dense_grad(p), ref) < RTOL, "factor %d" % k


@pytest.mark.parametrize("chain", [1, 0])
def test_sss_c1_chain_kernels_vs_oracle(built_lib, monkeypatch, chain):
    """BASELINE C1 (4096 -> 1000, 500 stages, statespace 16, batch 256) with the tcgen05 chain kernels forced -- the kernels
    bench.py times at 65 536 samples take over above 4 096 samples per GPU; below that the layer would use the SIMT scans."""
    monkeypatch.setenv("SNB200_SSS_PATH", "tc")
    monkeypatch.setenv("SNB200_SSS_TC_CHAIN", str(chain))
    monkeypatch.setenv("SNB200_SSS_TC_FUSED", "0")
    B = 256
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=1002))
    rng = np.random.default_rng(1002)
    X = rng.uniform(-1, 1, size=(B, 4096)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
    l32 = [[p.detach().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.sss_forward(torch.tensor(X), *l32, b, layer.dims_in, layer.dims_out); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    assert rms_rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    for li, name in enumerate("ABCDEFG"):
        got = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for p in getattr(layer, name)])
        ref = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)).reshape(-1) for p in l32[li]])
        assert rel_err(got, ref) < RTOL, "grad %s" % name
        assert rms_rel_err(got, ref) < RTOL, "grad %s (rms)" % name
    assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL


def test_sss_layer_with_more_chunks_than_the_scan_tables_vs_oracle(built_lib, monkeypatch):
    """4 400 stages of one input and one output = 275 chunks: more than the adjoint scan's shared-memory chunk table (256 entries, the
    rest is read from global memory) and too many for its per-CTA bias sums (the column-sum kernel beside the build-backward kernel
    takes over) -- the two paths the C1 / C5 layer (32 chunks) never takes."""
    monkeypatch.setenv("SNB200_SSS_PATH", "tc")
    B, n = 24, 4400
    layer = SSSLayer(n, n, 0.02, nb_states=n, initial_system_approx=random_mixed_system(n, n, n, 4, seed=1005))
    assert layer._use_tc_path()
    rng = np.random.default_rng(1005)
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32) / B
    l32 = [[p.detach().clone().requires_grad_(True) for p in getattr(layer, nm)] for nm in "ABCDEFG"]
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.sss_forward(torch.tensor(X), *l32, b, layer.dims_in, layer.dims_out); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    for li, name in enumerate("ABCDEFG"):
        got = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for p in getattr(layer, name)])
        ref = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)).reshape(-1) for p in l32[li]])
        if ref.size:
            assert rel_err(got, ref) < RTOL, "grad %s" % name
    assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL


@pytest.mark.parametrize("path", ["leaf", "dense"])
def test_hmat_c4h_shape_both_paths_vs_oracle(built_lib, monkeypatch, path):
    """BASELINE C4-H (2048 -> 1000, eta 0.5, min block 2: 2 560 leaves of rank <= 6) on the leaf-by-leaf kernels
    (sn_hmat_forward / sn_hmat_backward) and on the dense-block path."""
    from bench import HMatWorkload
    monkeypatch.setenv("SNB200_HMAT_PATH", path)
    wl = HMatWorkload()
    layer = HMatLayer(2048, 1000, 0.2, initial_hmatrix=wl.hmatrix())
    assert len(layer.hmatrix_components) == 2560 and layer._use_dense_path() == (path == "dense")
    B = 64
    rng = np.random.default_rng(4444)
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
    assert rms_rel_err(gl, gl_o) < RTOL and rms_rel_err(gr, gr_o) < RTOL
    assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL


def _ldr_case(n, share, B, seed, scale=1.0):
    np.random.seed(seed)
    layer = LDRLayer(n, n, share)
    if scale != 1.0:     # operators of larger norm: more Krylov terms matter
        with torch.no_grad():
            for k in (0, 1):
                p = layer.representation_matrices[k]
                layer.representation_matrices[k] = torch.nn.Parameter(torch.sparse_coo_tensor(p._indices(), p._values() * scale, p.shape))
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32) / B
    return layer, X, gy


@pytest.mark.parametrize("n,share,scale", [(512, 0.1, 1.0), (512, 0.1, 3.0), (2048, 0.1, 1.0)])
def test_ldr_config_size_vs_oracle(built_lib, n, share, scale):
    """LDR at n = 512 (r = 22) with the reference's Glorot-scale operators and with operators three times larger (slower decay
    of the Krylov series), and at BASELINE C4-L's n = 2048 (r = 99): against the oracle's float64 recurrence
    (oracle.layers_cpu.ldr_weight; the reference's literal matrix_power construction is O(r n^4 log n))."""
    B = 32
    layer, X, gy = _ldr_case(n, share, B, seed=700 + n, scale=scale)
    r = layer.representation_matrices[2].shape[1]
    assert r == {512: 22, 2048: 99}[n]
    rep = [p.detach().clone().requires_grad_(True) for p in layer.representation_matrices]
    b = layer.bias.detach().clone().requires_grad_(True)
    nb_terms = None
    if n > 512:          # all n powers cost minutes and gigabytes on the host; the dropped tail is below float64 resolution
        nb_terms = 48
        assert O.ldr_tail_bound(rep, nb_terms) < 1e-30
    yo = O.ldr_forward(torch.tensor(X), rep, b, (n, n), nb_terms=nb_terms); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    tol = 1e-4
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < tol
    A, Bm, G, H = layer.representation_matrices
    mA = rep[0].detach().to_dense().numpy() != 0; mB = rep[1].detach().to_dense().numpy() != 0
    gA_o = (rep[0].grad.to_dense() if rep[0].grad.is_sparse else rep[0].grad).numpy() * mA
    gB_o = (rep[1].grad.to_dense() if rep[1].grad.is_sparse else rep[1].grad).numpy() * mB
    assert A.grad.is_sparse and A.grad.dtype == torch.float64 and G.grad.dtype == torch.float64
    assert rel_err(dense_grad(A), gA_o) < tol and rel_err(dense_grad(Bm), gB_o) < tol
    assert rel_err(dense_grad(G), rep[2].grad.numpy()) < tol and rel_err(dense_grad(H), rep[3].grad.numpy()) < tol
    assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < tol
    assert 1 < layer.last_nb_terms < n


def test_tl_more_than_65535_krylov_rows(built_lib):
    """n * r = 512 * 128 = 65 536 rows in the stacked Krylov matrix (the row index used to sit on grid.y)."""
    rng = np.random.default_rng(6100)
    n, B = 512, 16
    layer = TLLayer(n, n, 0.5)
    r = layer.G.shape[1]
    assert n * r > 65535
    with torch.no_grad():
        layer.G.mul_(1.0 / np.sqrt(r)); layer.H.mul_(1.0 / np.sqrt(r))
    X = rng.uniform(-1, 1, size=(B, n)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n)).astype(np.float32) / B
    Gt, Ht, b = [t.detach().clone().requires_grad_(True) for t in (layer.G, layer.H, layer.bias)]
    yo = O.tl_forward(torch.tensor(X), Gt, Ht, b); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    y = layer(torch.tensor(X, device=DEV)); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), yo.detach().numpy()) < RTOL
    assert rel_err(dense_grad(layer.G), Gt.grad.numpy()) < 2 * RTOL and rel_err(dense_grad(layer.H), Ht.grad.numpy()) < 2 * RTOL


def test_input_gradients_of_the_layers_that_materialise_their_weight(built_lib, monkeypatch):
    """``grad_x`` is nullable in every backward; the reference gets it from autograd whenever the features
    require a gradient (a layer in the middle of a network).  TL, LDR, H-matrix (dense-block path) and PSM (dense-product path)
    return it from one more tensor-core GEMM; compared with autograd through the oracle."""
    import scipy.sparse
    from structurednets_b200.layers.psm_layer import PSMLayer
    rng = np.random.default_rng(9090)

    def check(layer, oracle_fn, n_in, n_out, B=24, tol=RTOL):
        X = rng.uniform(-1, 1, size=(B, n_in)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, n_out)).astype(np.float32)
        xo = torch.tensor(X, requires_grad=True)
        (oracle_fn(xo) * torch.tensor(gy)).sum().backward()
        lay = layer.to(DEV)
        xd = torch.tensor(X, device=DEV, requires_grad=True)
        (lay(xd) * torch.tensor(gy, device=DEV)).sum().backward()
        assert xd.grad is not None and rel_err(xd.grad.cpu().numpy(), xo.grad.numpy()) < tol

    n = 64
    tl = TLLayer(n, n, 0.3)
    check(tl, lambda x: O.tl_forward(x, tl.G.detach().cpu(), tl.H.detach().cpu(), tl.bias.detach().cpu()), n, n)
    np.random.seed(3)
    ldr = LDRLayer(n, n, 0.4)
    rep = [p.detach().clone() for p in ldr.representation_matrices]
    check(ldr, lambda x: O.ldr_forward(x, rep, ldr.bias.detach().cpu(), (n, n)), n, n, tol=1e-4)
    monkeypatch.setenv("SNB200_HMAT_PATH", "dense")
    hm = HMatLayer(64, 40, 0.6)
    comps = [(c.row_range.start, c.row_range.stop, c.col_range.start, c.col_range.stop, c.left_lr.detach().clone(), c.right_lr.detach().clone())
             for c in hm.hmatrix_components]
    check(hm, lambda x: O.hmat_forward(x, comps, hm.bias.detach().cpu(), 40), 64, 40)
    monkeypatch.setenv("SNB200_PSM_PATH", "dense")
    S = [scipy.sparse.random(40, 96, density=0.1, random_state=1, format="csr"), scipy.sparse.random(96, 96, density=0.1, random_state=2, format="csr"),
         scipy.sparse.random(96, 96, density=0.1, random_state=3, format="csr")]
    psm = PSMLayer(96, 40, sparse_matrices=S)
    dense = [torch.tensor(s.toarray()).float() for s in S]
    check(psm, lambda x: O.psm_forward(x, dense, psm.bias.detach().cpu()), 96, 40)


def test_structured_initialisation_on_the_gpu_matches_the_host(built_lib):
    """Hankel-SVD identification of an SSS layer and the per-leaf truncated SVDs of an H-matrix layer
    with the SVDs on the GPU (cuSOLVER through torch.linalg.svd, batched by leaf shape) give the same structured matrices as the
    numpy path (the factors themselves are unique only up to signs / a state-space similarity)."""
    from structurednets_b200.hmatrix.hmatrix import approximate_hmatrix
    from structurednets_b200.sss_identification import identify_mixed_system
    from structurednets_b200.synth import standard_dims
    rng = np.random.default_rng(8080)
    T = rng.uniform(-1, 1, size=(60, 96))
    di, do = standard_dims(96, 60, 12)
    host = identify_mixed_system(T, di, do, 4).to_matrix()
    dev = identify_mixed_system(T, di, do, 4, device="cuda").to_matrix()
    assert np.abs(host - dev).max() < 1e-9 * np.abs(host).max()
    W = rng.uniform(-1, 1, size=(200, 256))
    h_host = approximate_hmatrix(W, 0.3).to_dense_numpy()
    h_dev = approximate_hmatrix(W, 0.3, device="cuda").to_dense_numpy()
    assert np.abs(h_host - h_dev).max() < 1e-5 * np.abs(h_host).max()
    layer = HMatLayer(256, 200, 0.3, initial_weight_matrix=W, use_gpu=True)       # constructor path
    assert np.abs(layer.hmatrix.to_dense_numpy() - h_host).max() < 1e-5 * np.abs(h_host).max()
    sss = SSSLayer(96, 60, 0.5, nb_states=12, initial_weight_matrix=T, use_gpu=True)
    assert sss.statespace_dim > 0 and np.abs(sss.initial_weight_matrix - identify_mixed_system(T, di, do, sss.statespace_dim).to_matrix()).max() < 1e-8


@pytest.mark.parametrize("B", [64, 300])
def test_sss_input_gradient_vs_oracle(built_lib, B):
    """The tensor-core SSS path returns the gradient w.r.t. the input features (``grad_x`` nullable) from the
    adjoints its scans compute anyway: grad_u_j = [gy_j | lambda_{j+1} | mu_j] W_j per chunk.  C1's layer, against autograd through
    the oracle; the parameter gradients of the same backward must be unchanged."""
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=1003))
    rng = np.random.default_rng(1003)
    X = rng.uniform(-1, 1, size=(B, 4096)).astype(np.float32); gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
    l32 = [[p.detach().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]
    xo = torch.tensor(X, requires_grad=True)
    yo = O.sss_forward(xo, *l32, layer.bias.detach().clone(), layer.dims_in, layer.dims_out); (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(DEV)
    xd = torch.tensor(X, device=DEV, requires_grad=True)
    y = layer(xd); (y * torch.tensor(gy, device=DEV)).sum().backward()
    assert xd.grad is not None and xd.grad.shape == (B, 4096)
    assert rel_err(xd.grad.cpu().numpy(), xo.grad.numpy()) < RTOL
    assert rms_rel_err(xd.grad.cpu().numpy(), xo.grad.numpy()) < RTOL
    got = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for p in layer.B])
    ref = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)).reshape(-1) for p in l32[1]])
    assert rel_err(got, ref) < RTOL
