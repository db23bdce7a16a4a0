"""Sliced-ELL arrangement of the PSM sparsity patterns (host logic, no GPU)."""
import numpy as np
import pytest
import scipy.sparse

from structurednets_b200.layers.psm_layer import _sliced_ell


@pytest.mark.parametrize("shape,density,seed", [((50, 30), 0.2, 0), ((7, 100), 0.05, 1), ((130, 64), 0.5, 2), ((33, 33), 0.0, 3)])
def test_sliced_ell_reproduces_the_matrix(shape, density, seed):
    m = scipy.sparse.random(shape[0], shape[1], density=density, random_state=seed, format="coo")
    rows, cols, vals = m.row.astype(np.int64), m.col.astype(np.int64), m.data
    e = _sliced_ell(rows, cols, shape[0])
    assert e["slice_off"][0] == 0 and e["slice_off"][-1] == e["total"] and np.all(np.diff(e["slice_off"]) % 32 == 0)
    assert sorted(r for r in e["rowmap"] if r >= 0) == list(range(shape[0]))
    dense = np.zeros(shape)
    seen = np.zeros(len(vals), dtype=int)
    for s in range(e["nslices"]):
        base, end = e["slice_off"][s], e["slice_off"][s + 1]
        for p in range(base, end):
            if e["src"][p] >= 0:
                lane = (p - base) % 32
                r = e["rowmap"][s * 32 + lane]
                assert r == rows[e["src"][p]] and e["col"][p] == cols[e["src"][p]]
                dense[r, e["col"][p]] += vals[e["src"][p]]
                seen[e["src"][p]] += 1
    assert np.all(seen == 1)
    np.testing.assert_allclose(dense, m.toarray())
    # rows are sorted by length, so padding stays small
    if len(vals):
        assert e["total"] <= len(vals) + 32 * (e["nslices"] + shape[0] // 4 + 8)


def test_dense_product_patterns_match_scipy():
    """Host side of the PSM dense-product path: CSR / CSC / COO index arrays (with the map back into the parameter's own COO value
    array) for an uncoalesced pattern, against scipy."""
    import scipy.sparse
    from structurednets_b200.layers.psm_layer import PSMLayer, _dense_pattern
    m = scipy.sparse.random(37, 53, density=0.2, random_state=3, format="csr")
    layer = PSMLayer(53, 37, sparse_matrices=[m])
    p = layer.sparse_matrices[0]
    pat = _dense_pattern(p)
    vals = p._values().detach().numpy()
    rows, cols = p._indices()[0].numpy(), p._indices()[1].numpy()
    dense = np.zeros((37, 53), dtype=np.float64)
    np.add.at(dense, (rows, cols), vals)
    np.testing.assert_allclose(dense, m.toarray(), rtol=1e-6)
    for name, nmaj, major, minor in (("csr", 37, rows, cols), ("csc", 53, cols, rows)):
        ptr, idx, src = (pat[name + "_" + k].numpy() for k in ("ptr", "idx", "src"))
        assert ptr[0] == 0 and ptr[-1] == len(vals) and len(ptr) == nmaj + 1 and np.all(np.diff(ptr) >= 0)
        for r in range(nmaj):
            e = slice(ptr[r], ptr[r + 1])
            assert np.all(major[src[e]] == r)                      # every entry of the row maps back to a COO entry of that row
            np.testing.assert_array_equal(idx[e], minor[src[e]])
        assert sorted(src.tolist()) == list(range(len(vals)))     # a permutation of the COO entries
    np.testing.assert_array_equal(pat["coo_row"].numpy(), rows)
    np.testing.assert_array_equal(pat["coo_col"].numpy(), cols)
    assert layer.use_dense_path(4096) is False                    # 37 % 4 != 0: not eligible
    assert PSMLayer(53, 36, sparse_matrices=[scipy.sparse.random(36, 53, density=0.2, random_state=1, format="csr")]).use_dense_path(4096)
