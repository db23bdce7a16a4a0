"""TEST INFRASTRUCTURE ONLY -- torch (CPU) statement of the chunked formulation behind the tensor-core SSS path
(csrc/sss_tc.cu).  Differentiable, so autograd gives reference values for every intermediate the CUDA kernels
produce: chunk matrices W / scan coefficients SC, local GEMM output, chunk-boundary states, adjoints, dM.

Chunk j covers natural stages [k0, k1).  With u_j the chunk's input columns, s_j the causal state entering the chunk
(from below) and e_{j+1} the anticausal state entering it (from above):

    y_j     = T_j u_j + O_j s_j + O'_j e_{j+1}            T_j: lower triangle + D blocks (causal) + upper (anticausal)
    s_{j+1} = R_j u_j + Phi_j s_j
    e_j     = R'_j u_j + Phi'_j e_{j+1}

which is the stage recursion of reference layers/sss_layer.py:111-123 with the stages of a chunk multiplied out.
"""
import numpy as np
import torch

DS = 16      # state padding
PO = 32      # output rows per chunk (padded)
KB = 32      # input columns per k-block
KB_MAX = 5   # k-blocks per chunk
LMAX = 16    # stages per chunk


def make_chunks(dims_in, dims_out):
    """Greedy chunking: consecutive stages while out <= PO, in <= KB*KB_MAX, stages <= LMAX."""
    n = len(dims_in)
    chunks = []
    k = 0
    while k < n:
        k0 = k
        nin = nout = 0
        while k < n and k - k0 < LMAX and nin + int(dims_in[k]) <= KB * KB_MAX and nout + int(dims_out[k]) <= PO:
            nin += int(dims_in[k])
            nout += int(dims_out[k])
            k += 1
        assert k > k0, "a single stage exceeds the chunk limits"
        chunks.append((k0, k))
    return chunks


def chunk_matrices(A, B, C, D, E, F, G, dims_in, dims_out, k0, k1):
    """W (64 x KB*KB_MAX): rows 0..31 T, 32..47 R, 48..63 R'.  SC: Phi (16x16), Phi' (16x16), O (32x16), O' (32x16)."""
    f = A[0].dtype
    io = np.concatenate([[0], np.cumsum(dims_in)]).astype(int)
    oo = np.concatenate([[0], np.cumsum(dims_out)]).astype(int)
    c0, r0 = io[k0], oo[k0]
    m, p = io[k1] - c0, oo[k1] - r0
    W = torch.zeros((64, KB * KB_MAX), dtype=f)
    Phi = torch.zeros((DS, DS), dtype=f)
    Phip = torch.zeros((DS, DS), dtype=f)
    O = torch.zeros((PO, DS), dtype=f)
    Op = torch.zeros((PO, DS), dtype=f)
    Wrows, Orows, Oprows = [], [], []
    # causal: propagate the matrix "state as a function of [u_chunk, s_in]"
    d_in = A[k0].shape[1]
    Su = torch.zeros((d_in, m), dtype=f)          # state <- u
    Ss = torch.eye(d_in, dtype=f)                 # state <- s_in
    Tc = torch.zeros((p, m), dtype=f)
    Oc = torch.zeros((p, d_in), dtype=f)
    for k in range(k0, k1):
        rows = slice(oo[k] - r0, oo[k + 1] - r0)
        cols = slice(io[k] - c0, io[k + 1] - c0)
        Tk = C[k] @ Su
        Tk = torch.cat([Tk[:, :cols.start], Tk[:, cols] + D[k], Tk[:, cols.stop:]], dim=1)
        Tc = torch.cat([Tc[:rows.start], Tk, Tc[rows.stop:]], dim=0)
        Oc = torch.cat([Oc[:rows.start], C[k] @ Ss, Oc[rows.stop:]], dim=0)
        Bk = torch.zeros((B[k].shape[0], m), dtype=f)
        Bk = torch.cat([Bk[:, :cols.start], B[k], Bk[:, cols.stop:]], dim=1)
        Su = A[k] @ Su + Bk
        Ss = A[k] @ Ss
    Rc, Phic = Su, Ss                              # (d_out x m), (d_out x d_in)
    # anticausal: stages k1-1 .. k0
    e_in = E[k1 - 1].shape[1]
    Su = torch.zeros((e_in, m), dtype=f)
    Ss = torch.eye(e_in, dtype=f)
    Ta = torch.zeros((p, m), dtype=f)
    Oa = torch.zeros((p, e_in), dtype=f)
    for k in range(k1 - 1, k0 - 1, -1):
        rows = slice(oo[k] - r0, oo[k + 1] - r0)
        cols = slice(io[k] - c0, io[k + 1] - c0)
        Ta = torch.cat([Ta[:rows.start], G[k] @ Su, Ta[rows.stop:]], dim=0)
        Oa = torch.cat([Oa[:rows.start], G[k] @ Ss, Oa[rows.stop:]], dim=0)
        Fk = torch.zeros((F[k].shape[0], m), dtype=f)
        Fk = torch.cat([Fk[:, :cols.start], F[k], Fk[:, cols.stop:]], dim=1)
        Su = E[k] @ Su + Fk
        Ss = E[k] @ Ss
    Ra, Phia = Su, Ss
    pad = lambda M, r, c: torch.nn.functional.pad(M, (0, c - M.shape[1], 0, r - M.shape[0]))
    W = torch.cat([pad(Tc + Ta, PO, KB * KB_MAX), pad(Rc, DS, KB * KB_MAX), pad(Ra, DS, KB * KB_MAX)], dim=0)
    return W, pad(Phic, DS, DS), pad(Phia, DS, DS), pad(Oc, PO, DS), pad(Oa, PO, DS)


def forward_chunked(U, A, B, C, D, E, F, G, bias, dims_in, dims_out, chunks=None, return_all=False, mats=None):
    """Same result as oracle.layers_cpu.sss_forward through: local GEMM -> chunk scans -> fix-up."""
    if chunks is None:
        chunks = make_chunks(dims_in, dims_out)
    io = np.concatenate([[0], np.cumsum(dims_in)]).astype(int)
    oo = np.concatenate([[0], np.cumsum(dims_out)]).astype(int)
    Bn = U.shape[0]
    nc = len(chunks)
    if mats is None:
        mats = [chunk_matrices(A, B, C, D, E, F, G, dims_in, dims_out, k0, k1) for (k0, k1) in chunks]
    Upad = torch.nn.functional.pad(U, (0, KB * KB_MAX))
    loc = []
    for j, (k0, k1) in enumerate(chunks):
        W = mats[j][0]
        uj = Upad[:, io[k0]:io[k0] + KB * KB_MAX]
        # columns beyond the chunk hit zero weights
        loc.append(uj @ W.T)                       # (B x 64): y_local | r | r'
    s = [None] * (nc + 1)
    e = [None] * (nc + 1)
    s[0] = torch.zeros((Bn, DS), dtype=U.dtype)
    e[nc] = torch.zeros((Bn, DS), dtype=U.dtype)
    for j in range(nc - 1, -1, -1):
        e[j] = loc[j][:, 48:64] + e[j + 1] @ mats[j][2].T
    ys = []
    for j, (k0, k1) in enumerate(chunks):
        yj = loc[j][:, :32] + s[j] @ mats[j][3].T + e[j + 1] @ mats[j][4].T
        ys.append(yj[:, :oo[k1] - oo[k0]])
        s[j + 1] = loc[j][:, 32:48] + s[j] @ mats[j][1].T
    y = torch.cat(ys, dim=1)
    if bias is not None:
        y = y + bias
    if return_all:
        return y, dict(mats=mats, loc=loc, s=s, e=e, chunks=chunks)
    return y
