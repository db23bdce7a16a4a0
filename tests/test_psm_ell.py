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
