"""Deterministic synthetic structures for tests and benchmarks (host side, numpy only).

The reference builds its initial structures with third-party fitters that are not installed here
(tvsclib for SSS, pyfaust for PSM).  Every layer constructor accepts a ready-made
structure instead (``initial_system_approx=``, ``sparse_matrices=``, ``initial_hmatrix=`` ...); the
generators below produce such structures from a seed, with the shapes the reference would produce.
"""
import numpy as np


class _Stage:
    def __init__(self, A, B, C, D=None):
        self.A_matrix, self.B_matrix, self.C_matrix, self.D_matrix = A, B, C, D


class _StrictSystem:
    def __init__(self, stages):
        self.stages = stages


class SyntheticMixedSystem:
    """Duck-typed stand-in for ``tvsclib.mixed_system.MixedSystem`` with exactly the attributes
    ``SSSLayer`` reads (reference layers/sss_layer.py:45-46,70-80): ``dims_in``, ``dims_out``,
    ``to_matrix()``, ``causal_system.stages[k].{A,B,C,D}_matrix`` and
    ``anticausal_system.stages[k].{A,B,C}_matrix``.  Picklable."""

    def __init__(self, dims_in, dims_out, causal_stages, anticausal_stages):
        self.dims_in = np.asarray(dims_in, dtype="int32")
        self.dims_out = np.asarray(dims_out, dtype="int32")
        self.causal_system = _StrictSystem(causal_stages)
        self.anticausal_system = _StrictSystem(anticausal_stages)

    def to_matrix(self) -> np.ndarray:
        """Dense matrix of the system: one sweep per direction carrying the states of ALL earlier (later) input columns at once, so the
        cost is n small matrix products of growing width instead of n^2 Python iterations."""
        n = len(self.dims_in)
        io = np.concatenate([[0], np.cumsum(self.dims_in)]).astype(int)
        oo = np.concatenate([[0], np.cumsum(self.dims_out)]).astype(int)
        T = np.zeros((oo[-1], io[-1]))
        cs, acs = self.causal_system.stages, self.anticausal_system.stages
        acc = np.zeros((cs[0].A_matrix.shape[1], 0))          # states (entering stage i) driven by the columns of stages < i
        for i in range(n):
            T[oo[i]:oo[i + 1], io[i]:io[i + 1]] = cs[i].D_matrix
            if acc.shape[1]:
                T[oo[i]:oo[i + 1], :io[i]] = cs[i].C_matrix @ acc
            acc = np.concatenate([cs[i].A_matrix @ acc, cs[i].B_matrix], axis=1)
        acc = np.zeros((acs[n - 1].A_matrix.shape[1], 0))     # anticausal: columns of stages > i, in column order
        for i in range(n - 1, -1, -1):
            if acc.shape[1]:
                T[oo[i]:oo[i + 1], io[i + 1]:] = acs[i].C_matrix @ acc
            acc = np.concatenate([acs[i].B_matrix, acs[i].A_matrix @ acc], axis=1)
        return T


def standard_dims(input_size: int, output_size: int, nb_states: int):
    dims_in = int(input_size / nb_states) * np.ones((nb_states,), dtype="int32")
    dims_in[:(input_size - np.sum(dims_in))] += 1
    dims_out = int(output_size / nb_states) * np.ones((nb_states,), dtype="int32")
    dims_out[:(output_size - np.sum(dims_out))] += 1
    return dims_in, dims_out


def random_mixed_system(input_dim: int, output_dim: int, nb_states: int, statespace_dim: int, seed: int = 0,
                        ragged_state_dims: bool = False) -> SyntheticMixedSystem:
    """Random stable mixed system with the boundary shapes tvsclib produces:
    causal state dims d_0 = d_n = 0, anticausal state dims e_0 = e_n = 0, interior = statespace_dim
    (or smaller near the ends when ``ragged_state_dims``).  Entries are U(-1,1) * 0.5 / sqrt(cols) so
    the recursions neither blow up nor vanish."""
    rng = np.random.default_rng(seed)
    dims_in, dims_out = standard_dims(input_dim, output_dim, nb_states)
    n = nb_states
    d = np.full(n + 1, statespace_dim, dtype=int)
    if ragged_state_dims:
        for k in range(n + 1):
            d[k] = min(statespace_dim, 1 + 2 * min(k, n - k))
    d[0] = 0
    d[n] = 0
    e = d.copy()

    def mat(r, c):
        return (rng.uniform(-1, 1, size=(r, c)) * 0.5 / np.sqrt(max(c, 1))).astype(np.float64)

    causal, anti = [], []
    for k in range(n):
        causal.append(_Stage(mat(d[k + 1], d[k]), mat(d[k + 1], dims_in[k]), mat(dims_out[k], d[k]), mat(dims_out[k], dims_in[k])))
        # anticausal state z_k has dim e[k]; stage k maps z_{k+1} (dim e[k+1]) -> z_k (dim e[k])
        anti.append(_Stage(mat(e[k], e[k + 1]), mat(e[k], dims_in[k]), mat(dims_out[k], e[k + 1])))
    return SyntheticMixedSystem(dims_in, dims_out, causal, anti)
