"""TEST INFRASTRUCTURE ONLY -- restatement of the reference layers' forward passes with the reference's own torch ops
(CPU tensors in the tests and the CPU baseline; bench.py's ``gpu_aten_baseline`` leg feeds the same functions CUDA tensors, which
is what the reference's ``use_gpu=True`` does: the same ATen ops on the device).

Every function here follows one ``forward`` of ``/root/reference/src/structurednets/layers``
op-for-op with torch CPU ops, so torch autograd supplies exactly the backward the
reference gets (the reference has no hand-written backward: SURVEY.md section 8a, rows a4/a7).
They take plain tensors (no nn.Module), so the same seeded parameters can be fed to the
oracle and to the CUDA path.  Pinned against the unmodified reference modules by
``tests/golden/make_golden.py`` -> ``tests/golden/*.npz`` -> ``tests/test_oracle_golden.py``.

Nothing in ``structurednets_b200`` imports this file.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# SSS  (reference: layers/sss_layer.py:99-131)
# --------------------------------------------------------------------------------------
def sss_forward(U: torch.Tensor, A: Sequence[torch.Tensor], B: Sequence[torch.Tensor],
                C: Sequence[torch.Tensor], D: Sequence[torch.Tensor], E: Sequence[torch.Tensor],
                F: Sequence[torch.Tensor], G: Sequence[torch.Tensor], bias: Optional[torch.Tensor],
                dims_in: Sequence[int], dims_out: Sequence[int]) -> torch.Tensor:
    """Causal sweep k=0..n-1 interleaved with the anticausal sweep j=n-1..0, exactly as the
    reference loop ``sss_layer.py:111-123``; transpose + bias as ``:125-127``."""
    n = len(dims_in)
    in_off = np.concatenate([[0], np.cumsum(dims_in)]).astype(int)    # sss_layer.py:93-94
    out_off = np.concatenate([[0], np.cumsum(dims_out)]).astype(int)  # sss_layer.py:96-97
    Ut = torch.transpose(U, 1, 0)
    dev = U.device                                                    # sss_layer.py:106-109 (`.to(device)` when use_gpu)
    y = torch.zeros((int(out_off[-1]), U.shape[0]), dtype=U.dtype, device=dev)    # sss_layer.py:101
    x = torch.zeros((0, U.shape[0]), dtype=U.dtype, device=dev)                   # :103
    z = torch.zeros((0, U.shape[0]), dtype=U.dtype, device=dev)                   # :104
    for k in range(n):
        u_k = Ut[in_off[k]:in_off[k + 1], :]                                              # :114
        y[out_off[k]:out_off[k + 1], :] += torch.matmul(C[k], x) + torch.matmul(D[k], u_k)  # :115
        x = torch.matmul(A[k], x) + torch.matmul(B[k], u_k)                               # :116
        j = n - 1 - k                                                                     # :118
        u_j = Ut[in_off[j]:in_off[j + 1], :]                                              # :121
        y[out_off[j]:out_off[j + 1], :] += torch.matmul(G[j], z)                          # :122
        z = torch.matmul(E[j], z) + torch.matmul(F[j], u_j)                               # :123
    y = torch.transpose(y, 1, 0)                                                          # :125
    if bias is not None:
        y = y + bias                                                                      # :126-127
    return y


def sss_to_matrix(A, B, C, D, E, F, G, dims_in, dims_out) -> np.ndarray:
    """Dense (out, in) matrix of the mixed system in float64: block (i,j) is D_i if i==j,
    C_i A_{i-1}..A_{j+1} B_j if i>j, G_i E_{i+1}..E_{j-1} F_j if i<j.  This is
    what tvsclib's ``MixedSystem.to_matrix()`` returns and what the reference's own
    ``test_sss_layer`` compares the layer against (tests/test_layers.py:118-121)."""
    n = len(dims_in)
    in_off = np.concatenate([[0], np.cumsum(dims_in)]).astype(int)
    out_off = np.concatenate([[0], np.cumsum(dims_out)]).astype(int)
    f64 = lambda t: np.asarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float64)
    A, B, C, D, E, F, G = [[f64(m) for m in lst] for lst in (A, B, C, D, E, F, G)]
    T = np.zeros((int(out_off[-1]), int(in_off[-1])))
    for j in range(n):
        T[out_off[j]:out_off[j + 1], in_off[j]:in_off[j + 1]] = D[j]
        acc = B[j]                       # state contribution of u_j after stage j
        for i in range(j + 1, n):
            T[out_off[i]:out_off[i + 1], in_off[j]:in_off[j + 1]] = C[i] @ acc
            acc = A[i] @ acc
        acc = F[j]
        for i in range(j - 1, -1, -1):
            T[out_off[i]:out_off[i + 1], in_off[j]:in_off[j + 1]] = G[i] @ acc
            acc = E[i] @ acc
    return T


# --------------------------------------------------------------------------------------
# PSM  (reference: layers/psm_layer.py:47-60 and :36-45)
# --------------------------------------------------------------------------------------
def psm_forward_literal(U: torch.Tensor, factors: Sequence[torch.Tensor], bias: Optional[torch.Tensor]) -> torch.Tensor:
    """The literal reference order ``psm_layer.py:51-54``: last factor first, then factors
    ``[0], [1], ...`` -- correct only for <= 2 factors (SURVEY.md finding F2).  ``factors`` are dense
    or sparse-COO tensors; sparse ones are densified like ``:51``."""
    dense = [f.to_dense() if f.is_sparse else f for f in factors]
    y = torch.mm(dense[-1], U.T)
    for m in dense[:-1]:
        y = torch.mm(m, y)
    y = y.T
    if bias is not None:
        y = y + bias
    return y


def psm_forward(U: torch.Tensor, factors: Sequence[torch.Tensor], bias: Optional[torch.Tensor]) -> torch.Tensor:
    """The intended chain W = S0 S1 ... S(n-1) (psm_approximator.py:96-105; psm_layer.py:76-78):
    ``psm_layer.py:51-58`` with the loop over ``reversed(dense[:-1])``.  Identical to the literal
    order for 1 or 2 factors."""
    dense = [f.to_dense() if f.is_sparse else f for f in factors]
    y = torch.mm(dense[-1], U.T)
    for m in reversed(dense[:-1]):
        y = torch.mm(m, y)
    y = y.T
    if bias is not None:
        y = y + bias
    return y


# --------------------------------------------------------------------------------------
# LR  (reference: layers/lr_layer.py:38-46)
# --------------------------------------------------------------------------------------
def lr_forward(U, left_lr, right_lr, bias):
    y = torch.matmul(right_lr, U.T)      # lr_layer.py:39
    y = torch.matmul(left_lr, y)         # :40
    y = y.T                              # :41
    if bias is not None:
        y = y + bias                     # :43-44
    return y


# --------------------------------------------------------------------------------------
# H-matrix  (reference: layers/hmat_layer.py:34-49)
# --------------------------------------------------------------------------------------
def hmat_forward(U, components: Sequence[Tuple[int, int, int, int, torch.Tensor, torch.Tensor]], bias, output_dim: int):
    """``components``: (row_start, row_stop, col_start, col_stop, left_lr, right_lr) per leaf with
    rank > 0, in the reference's traversal order (hmat_layer.py:26-28)."""
    res = torch.zeros((output_dim, U.shape[0]), dtype=U.dtype, device=U.device)  # :35-37 (`.to(device)` when use_gpu)
    U_T = U.T                                                                   # :39
    for (r0, r1, c0, c1, L, R) in components:
        res[r0:r1, :] += torch.matmul(L, torch.matmul(R, U_T[c0:c1, :]))        # :43
    res = res.T                                                                 # :44
    if bias is not None:
        res = res + bias                                                        # :46-47
    return res


def hmat_to_dense(components, shape) -> np.ndarray:
    """hmatrix/hmatrix.py:13-21 in float64."""
    W = np.zeros(shape)
    for (r0, r1, c0, c1, L, R) in components:
        W[r0:r1, c0:c1] = L.detach().double().numpy() @ R.detach().double().numpy()
    return W


# --------------------------------------------------------------------------------------
# LDR  (reference: approximators/ldr_approximator.py:29-39, layers/ldr_layer.py:44-61)
# --------------------------------------------------------------------------------------
def ldr_weight_literal(rep: Sequence[torch.Tensor], shape) -> torch.Tensor:
    """Op-for-op ``build_weight_matrix_torch`` (matrix_power per column) -- O(r n^4 log n); only
    usable for n <= ~64.  Used to validate :func:`ldr_weight`."""
    res = torch.zeros(shape)
    n = shape[0]
    Ad = rep[0].to_dense() if rep[0].is_sparse else rep[0]
    Bd = rep[1].to_dense() if rep[1].is_sparse else rep[1]
    BdT = torch.transpose(Bd, 0, 1)
    for i in range(rep[2].shape[1]):
        g = torch.index_select(rep[2], 1, torch.LongTensor([i]))
        h = torch.index_select(rep[3], 1, torch.LongTensor([i]))
        ka = torch.cat([g] + [torch.matmul(torch.matrix_power(Ad, j), g) for j in range(1, n)], 1)
        kb = torch.cat([h] + [torch.matmul(torch.matrix_power(BdT, j), h) for j in range(1, n)], 1)
        res += torch.matmul(ka, kb.T)
    return res


def ldr_tail_bound(rep: Sequence[torch.Tensor], nb_terms: int) -> float:
    """Upper bound of everything ``ldr_weight(..., nb_terms=nb_terms)`` leaves out, relative to the j = 0 term's bound:
    sum_{j >= nb_terms} (|A|_inf |B|_1)^j -- meaningful only when the operators' norms are below 1."""
    Ad = (rep[0].to_dense() if rep[0].is_sparse else rep[0]).detach().abs()
    Bd = (rep[1].to_dense() if rep[1].is_sparse else rep[1]).detach().abs()
    q = float(Ad.sum(1).max() * Ad.sum(0).max()) ** 0.5 * float(Bd.sum(1).max() * Bd.sum(0).max()) ** 0.5
    return float("inf") if q >= 1 else q ** nb_terms / (1 - q)


def ldr_weight(rep: Sequence[torch.Tensor], shape, nb_terms: Optional[int] = None) -> torch.Tensor:
    """Same matrix as ``ldr_approximator.py:29-39`` with the Krylov matrices built by the
    recurrence k_{j+1} = M k_j (as the reference's own numpy TL code does,
    approximators/tl_approximator.py:14-19) in float64, accumulated into float32 per term like
    ``:31,37``.  Differentiable (autograd through the recurrence).  ``nb_terms`` (tests at n = 2048 only): keep the
    first nb_terms of the n Krylov powers; the caller asserts with :func:`ldr_tail_bound` that the dropped tail is below
    float64 resolution, so the matrix is the same."""
    n = shape[0] if nb_terms is None else min(shape[0], nb_terms)
    Ad = rep[0].to_dense() if rep[0].is_sparse else rep[0]
    Bd = rep[1].to_dense() if rep[1].is_sparse else rep[1]
    BdT = torch.transpose(Bd, 0, 1)
    Gm, Hm = rep[2], rep[3]
    ka = [Gm]                      # all displacement-rank columns at once: (n, r) per power
    kb = [Hm]
    for _ in range(1, n):
        ka.append(torch.matmul(Ad, ka[-1]))
        kb.append(torch.matmul(BdT, kb[-1]))
    KA = torch.stack(ka, 2)        # (n, r, n): KA[:, i, j] = A^j g_i
    KB = torch.stack(kb, 2)
    W = torch.einsum("pij,qij->pq", KA, KB)
    return W.float()


def ldr_forward(U, rep, bias, shape, literal=False, nb_terms=None):
    W = ldr_weight_literal(rep, shape) if literal else ldr_weight(rep, shape, nb_terms=nb_terms)
    y = torch.matmul(W, U.T).T           # ldr_layer.py:54-55
    if bias is not None:
        y = y + bias                     # :57-58
    return y


# --------------------------------------------------------------------------------------
# Toeplitz-like  (reference: layers/tl_layer.py:11-18, 51-68)
# --------------------------------------------------------------------------------------
def _zf_krylov(f: float, vec: torch.Tensor) -> torch.Tensor:
    """tl_layer.py:11-18 (column i = roll(vec, i), first i entries scaled by f), built without
    the in-place writes so autograd stays cheap: K[p, i] = vec[(p - i) mod n] * (f if p < i else 1)."""
    n = vec.shape[0]
    p = torch.arange(n).unsqueeze(1)
    i = torch.arange(n).unsqueeze(0)
    K = vec[(p - i) % n]
    scale = torch.where(p < i, torch.tensor(float(f), dtype=vec.dtype), torch.tensor(1.0, dtype=vec.dtype))
    return K * scale


def tl_weight(Gm: torch.Tensor, Hm: torch.Tensor) -> torch.Tensor:
    r = Gm.shape[1]
    return 0.5 * sum(_zf_krylov(1, Gm[:, j]) @ _zf_krylov(-1, torch.flipud(Hm[j, :])) for j in range(r))  # tl_layer.py:59


def tl_forward(U, Gm, Hm, bias):
    if Gm.shape[1] > 0:
        y = torch.matmul(tl_weight(Gm, Hm), U.T).T      # tl_layer.py:59-61
    else:
        y = torch.zeros((U.shape[0], Gm.shape[0]))      # :62-63
    if bias is not None:
        y = y + bias                                    # :65-66
    return y


# --------------------------------------------------------------------------------------
# loss used by every parity test and by the CPU baseline: mean-square of the output, i.e. the
# ``y.square().mean().backward()`` step BASELINE.md section 2 timed.
# --------------------------------------------------------------------------------------
def mse_to_ones(y: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.mse_loss(y, torch.ones_like(y))
