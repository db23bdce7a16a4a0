"""numpy statement of the LDR formulation of csrc/ldr_tl.cu, kernel by kernel (band gather, Krylov recurrences with column
max-norms, device-side series length, stacked contraction, adjoint recurrences with the per-column power counts): the CPU
tests check it against the oracle's autograd, the GPU tests check the kernels against the oracle.  Test infrastructure only."""
import numpy as np


def band_of(idx, vals, n):
    """lo | di | up | corner(0,n-1), corner(n-1,0) -- layers/ldr_layer.py band_slots + band_gather_kernel."""
    band = np.zeros(3 * n + 2)
    for (r, c), v in zip(idx.T, vals):
        if c == r:
            band[n + r] += v
        elif c == r + 1:
            band[2 * n + r] += v
        elif c == r - 1:
            band[r] += v
        elif n > 2 and r == 0 and c == n - 1:
            band[3 * n] += v
        elif n > 2 and r == n - 1 and c == 0:
            band[3 * n + 1] += v
        else:
            raise ValueError("entry outside the tridiagonal-plus-corners pattern")
    return band


def band_transpose(b, n):
    t = np.zeros_like(b)
    t[1:n] = b[2 * n:3 * n - 1]
    t[n:2 * n] = b[n:2 * n]
    t[2 * n:3 * n - 1] = b[1:n]
    t[3 * n], t[3 * n + 1] = b[3 * n + 1], b[3 * n]
    return t


def band_apply(band, v, n):
    """M v for every column of v (n x c)."""
    lo, di, up = band[:n], band[n:2 * n], band[2 * n:3 * n]
    w = di[:, None] * v
    w[1:] += lo[1:, None] * v[:-1]
    w[:-1] += up[:-1, None] * v[1:]
    if n > 2:
        w[0] += band[3 * n] * v[n - 1]
        w[n - 1] += band[3 * n + 1] * v[0]
    return w


def band_apply_T(band, a, n):
    lo, di, up = band[:n], band[n:2 * n], band[2 * n:3 * n]
    w = di[:, None] * a
    w[:-1] += lo[1:, None] * a[1:]
    w[1:] += up[:-1, None] * a[:-1]
    if n > 2:
        w[n - 1] += band[3 * n] * a[0]
        w[0] += band[3 * n + 1] * a[n - 1]
    return w


def forward(bandA, bandBt, G, H, J, rel_tol):
    n, r = G.shape
    K = [np.zeros((J, n, r)), np.zeros((J, n, r))]
    norms = np.zeros((2, J, r))
    for op, (band, src) in enumerate(((bandA, G), (bandBt, H))):
        v = src.copy()
        for j in range(J):
            K[op][j] = v
            norms[op, j] = np.abs(v).max(axis=0)
            v = band_apply(band, v, n)
    bound = (norms[0] * norms[1]).sum(axis=1)
    terms, conv, mx = J, J >= n, 0.0
    for j in range(J):
        if j > 0 and not (bound[j] > rel_tol * mx):
            terms, conv = j, True
            break
        mx = max(mx, bound[j])
    rows = max(32, (J * r + 31) // 32 * 32)
    k_eff = min((terms * r + 31) // 32 * 32, rows)
    KA = np.zeros((rows, n)); KB = np.zeros((rows, n))
    KA[:J * r] = K[0].transpose(0, 2, 1).reshape(J * r, n)      # row j r + i = column i of power j
    KB[:J * r] = K[1].transpose(0, 2, 1).reshape(J * r, n)
    W = KA[:k_eff].astype(np.float32).astype(np.float64).T @ KB[:k_eff].astype(np.float32).astype(np.float64)
    return W, dict(K=K, KA=KA, KB=KB, terms=terms, k_eff=k_eff, conv=conv, rows=rows)


def backward(bandA, bandBt, info, dW, J):
    K, KA, KB, k_eff = info["K"], info["KA"], info["KB"], info["k_eff"]
    n, r = K[0].shape[1:]
    dKA = KB.astype(np.float32).astype(np.float64) @ dW.T          # [k][p] = sum_q KB[k][q] dW[p][q]
    dKB = KA.astype(np.float32).astype(np.float64) @ dW            # [k][q] = sum_p KA[k][p] dW[p][q]
    gbands, gsrc = [], []
    for op, (band, dK) in enumerate(((bandA, dKA), (bandBt, dKB))):
        gb = np.zeros(3 * n + 2)
        a = np.zeros((n, r))
        for j in range(J - 1, -1, -1):
            live = (j * r + np.arange(r)) < k_eff
            kj = K[op][j]
            gb[n:2 * n] += (a * kj).sum(axis=1)
            gb[1:n] += (a[1:] * kj[:-1]).sum(axis=1)
            gb[2 * n:3 * n - 1] += (a[:-1] * kj[1:]).sum(axis=1)
            if n > 2:
                gb[3 * n] += (a[0] * kj[n - 1]).sum()
                gb[3 * n + 1] += (a[n - 1] * kj[0]).sum()
            a = np.where(live[None, :], band_apply_T(band, a, n) + dK[j * r:(j + 1) * r].T if (j + 1) * r <= dK.shape[0] else 0.0, 0.0)
        gbands.append(gb)
        gsrc.append(a)
    return gbands[0], band_transpose(gbands[1], n), gsrc[0], gsrc[1]
