"""Time-varying system identification of a dense matrix by Hankel-matrix SVDs (host side, numpy): the initial values of an
``SSSLayer`` built from ``initial_weight_matrix=`` when tvsclib is not installed.

The reference delegates this to tvsclib (``SystemIdentificationSVD``, reference layers/sss_layer.py:65-68), an unpinned git-only
dependency; its result is unique only up to a state-space similarity and no reference test pins its values (SURVEY.md section 8c:
"SSS-from-dense initialisation: parity unpinned").  What the reference's tests do pin is self-consistency -- the layer's forward equals
``to_matrix()`` of the system (tests/test_layers.py:118-121) -- and that is what this realisation is tested for, together with exact
recovery when the state dimension is not truncated and a monotone approximation error when it is.

For a block partition (dims_in, dims_out) with n stages:
    causal part      x_{k+1} = A_k x_k + B_k u_k ,   y_k += C_k x_k + D_k u_k          T[i][j] = C_i A_{i-1} .. A_{j+1} B_j   (i > j)
    anticausal part  x'_k = E_k x'_{k+1} + F_k u_k ,  y_k += G_k x'_{k+1}               T[i][j] = G_i E_{i+1} .. E_{j-1} F_j   (i < j)
The causal state entering stage k factors the Hankel block H_k = T[rows of stages >= k, columns of stages < k] = O_k R_k (truncated
SVD, both factors carrying the square root of the singular values); C_k is the first block row of O_k, [A_k R_k | B_k] = R_{k+1}.
The anticausal part is the mirror image on H'_k = T[rows of stages <= k, columns of stages > k].
"""
import numpy as np

from structurednets_b200.synth import SyntheticMixedSystem, _Stage

_SVD_DEVICE = None     # set by identify_mixed_system(device=...): the Hankel SVDs then run in cuSOLVER through torch.linalg.svd


def _svd(H: np.ndarray):
    if _SVD_DEVICE is None:
        return np.linalg.svd(H, full_matrices=False)
    import torch
    U, S, Vt = torch.linalg.svd(torch.as_tensor(H, dtype=torch.float64, device=_SVD_DEVICE), full_matrices=False)
    return U.cpu().numpy(), S.cpu().numpy(), Vt.cpu().numpy()


def _factor(H: np.ndarray, max_states: int, rel_tol: float):
    """H ~ O R with O = U sqrt(S), R = sqrt(S) V^T, at most max_states columns; also returns pinv(R) = V / sqrt(S)."""
    if H.shape[0] == 0 or H.shape[1] == 0 or max_states <= 0:
        return np.zeros((H.shape[0], 0)), np.zeros((0, H.shape[1])), np.zeros((H.shape[1], 0))
    U, S, Vt = _svd(H)
    d = int(min(max_states, np.sum(S > rel_tol * max(S[0], 1e-300))))
    rs = np.sqrt(S[:d])
    return U[:, :d] * rs, rs[:, None] * Vt[:d], Vt[:d].T / rs


def identify_mixed_system(T, dims_in, dims_out, max_states: int, rel_tol: float = 1e-12, device=None) -> SyntheticMixedSystem:
    """``device`` (e.g. "cuda"): run the 2 (n - 1) Hankel-block SVDs on that device in float64 (cuSOLVER via torch.linalg.svd) --
    SURVEY.md section 8f rank 3, "GPU-side structured initialisation"; the realisation is the same up to the SVD's sign choices."""
    global _SVD_DEVICE
    prev, _SVD_DEVICE = _SVD_DEVICE, device
    try:
        return _identify(T, dims_in, dims_out, max_states, rel_tol)
    finally:
        _SVD_DEVICE = prev


def _identify(T, dims_in, dims_out, max_states: int, rel_tol: float) -> SyntheticMixedSystem:
    T = np.asarray(T, dtype=np.float64)
    dims_in, dims_out = np.asarray(dims_in, dtype=int), np.asarray(dims_out, dtype=int)
    n = len(dims_in)
    assert len(dims_out) == n and T.shape == (int(dims_out.sum()), int(dims_in.sum())), "the block partition does not match the matrix"
    io = np.concatenate([[0], np.cumsum(dims_in)])
    oo = np.concatenate([[0], np.cumsum(dims_out)])

    # causal: factors of H_k for k = 0 .. n (H_0 and H_n are empty: no state before the first / after the last stage)
    fac = [_factor(T[oo[k]:, :io[k]], max_states if 0 < k < n else 0, rel_tol) for k in range(n + 1)]
    causal = []
    for k in range(n):
        O_k, R_k, Rk_pinv = fac[k]
        R_next = fac[k + 1][1]
        A = R_next[:, :io[k]] @ Rk_pinv                      # (d_{k+1} x d_k)
        B = R_next[:, io[k]:io[k + 1]]                       # (d_{k+1} x in_k)
        C = O_k[:dims_out[k], :]                             # (out_k x d_k)
        D = T[oo[k]:oo[k + 1], io[k]:io[k + 1]].copy()
        causal.append(_Stage(A, B, C, D))

    # anticausal: x'_{k+1} (entering stage k from the right) factors H'_k = T[:oo[k+1], io[k+1]:]; index m = k + 1 below
    fac = [None] * (n + 1)
    fac[0] = _factor(np.zeros((0, T.shape[1])), 0, rel_tol)   # nothing leaves stage 0 to the left
    for m in range(1, n + 1):
        fac[m] = _factor(T[:oo[m], io[m]:], max_states if m < n else 0, rel_tol)
    anticausal = []
    for k in range(n):
        _, R_k, _ = fac[k]                                   # (e_k x inputs of stages >= k)
        O_next, _, Rnext_pinv = fac[k + 1]
        F = R_k[:, :dims_in[k]]                              # (e_k x in_k)
        E = R_k[:, dims_in[k]:] @ Rnext_pinv                 # (e_k x e_{k+1})
        G = O_next[oo[k]:oo[k + 1], :]                       # (out_k x e_{k+1})
        anticausal.append(_Stage(E, F, G))
    return SyntheticMixedSystem(dims_in, dims_out, causal, anticausal)
