"""numpy emulation of what csrc/sss.cu does with the stage / chunk tables (test helper).

It executes the *plan* (sn_sss_stage / sn_sss_chunk rows produced by SSSLayer.build_host_plan) the way
the kernels do -- pack with state rows first, chunked sweeps, first/second visit of y, checkpoints,
recompute + adjoint + outer-product gradients scattered to the natural parameter offsets -- so the host
logic can be checked against the oracle on the CPU.
"""
import numpy as np


def pack(stages, meta, flat):
    RP, KP = meta["rows_pad"], meta["k_pad"]
    out = {}
    for d in range(2):
        for kk in range(stages.shape[1]):
            in_off, in_dim, out_off, out_dim, d_in, d_out, oys, oyu, oss, osu, poff, k = stages[d, kk, :12]
            P = np.zeros((RP, KP), dtype=np.float32)
            P[:d_out, :d_in] = flat[oss:oss + d_out * d_in].reshape(d_out, d_in)
            P[:d_out, d_in:d_in + in_dim] = flat[osu:osu + d_out * in_dim].reshape(d_out, in_dim)
            P[d_out:d_out + out_dim, :d_in] = flat[oys:oys + out_dim * d_in].reshape(out_dim, d_in)
            if oyu >= 0:
                P[d_out:d_out + out_dim, d_in:d_in + in_dim] = flat[oyu:oyu + out_dim * in_dim].reshape(out_dim, in_dim)
            out[(d, kk)] = P
    return out


def forward(stages, chunks, meta, flat, X, bias):
    B = X.shape[0]
    P = pack(stages, meta, flat)
    y = np.full((B, meta["output_dim"]), np.nan, dtype=np.float32)
    ckpt = {}
    order = []  # emulate "all first-visit chunks of both directions, then second-visit chunks"
    for second in (0, 1):
        for d in range(2):
            order += [(d, ch) for ch in range(chunks.shape[1]) if chunks[d, ch, 6] == second]
    state = {0: np.zeros((0, B), np.float32), 1: np.zeros((0, B), np.float32)}
    for (d, ch) in order:
        kb, ke, col0, ncols, row0, nrows, second = chunks[d, ch][:7]
        ckpt[(d, ch)] = state[d].copy()
        yc = np.zeros((nrows, B), np.float32)
        for kk in range(kb, ke):
            in_off, in_dim, out_off, out_dim, d_in, d_out = stages[d, kk, :6]
            assert state[d].shape[0] == d_in
            inp = np.concatenate([state[d], X[:, in_off:in_off + in_dim].T], 0)
            res = P[(d, kk)][:d_out + out_dim, :d_in + in_dim] @ inp
            state[d] = res[:d_out]
            yc[out_off - row0:out_off - row0 + out_dim] = res[d_out:]
        if second:
            y[:, row0:row0 + nrows] += yc.T
        else:
            y[:, row0:row0 + nrows] = yc.T + (bias[row0:row0 + nrows] if bias is not None else 0)
    return y, ckpt


def backward(stages, chunks, meta, flat, X, gy, ckpt):
    B = X.shape[0]
    P = pack(stages, meta, flat)
    g = np.zeros_like(flat)
    for d in range(2):
        lcar = np.zeros((0, B), np.float32)
        for ch in range(chunks.shape[1] - 1, -1, -1):
            kb, ke, col0, ncols, row0, nrows, second = chunks[d, ch][:7]
            xs = [ckpt[(d, ch)]]
            for kk in range(kb, ke - 1):
                in_off, in_dim, out_off, out_dim, d_in, d_out = stages[d, kk, :6]
                inp = np.concatenate([xs[-1], X[:, in_off:in_off + in_dim].T], 0)
                xs.append(P[(d, kk)][:d_out, :d_in + in_dim] @ inp)
            lout = lcar
            for kk in range(ke - 1, kb - 1, -1):
                in_off, in_dim, out_off, out_dim, d_in, d_out, oys, oyu, oss, osu = stages[d, kk, :10]
                gout = np.concatenate([lout, gy[:, out_off:out_off + out_dim].T], 0)   # (d_out+out_dim, B)
                inp = np.concatenate([xs[kk - kb], X[:, in_off:in_off + in_dim].T], 0)  # (d_in+in_dim, B)
                dP = gout @ inp.T
                g[oss:oss + d_out * d_in] += dP[:d_out, :d_in].reshape(-1)
                g[osu:osu + d_out * in_dim] += dP[:d_out, d_in:].reshape(-1)
                g[oys:oys + out_dim * d_in] += dP[d_out:, :d_in].reshape(-1)
                if oyu >= 0:
                    g[oyu:oyu + out_dim * in_dim] += dP[d_out:, d_in:].reshape(-1)
                lout = P[(d, kk)][:d_out + out_dim, :d_in].T @ gout
            lcar = lout
    return g, gy.sum(0)
