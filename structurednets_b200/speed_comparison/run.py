"""B200 counterpart of the reference's only native component, ``speed_comparison/run.c`` (reference :123-368): reads the same
``layer_description.txt`` (written by ``prepare_run.py:33-82``), rebuilds the dense layer and the SSS layer from it, checks both
against the file's checksum vectors with run.c's protocol (absolute error <= 1e-3 in every output dimension, run.c:177,357, exit
code 1 and an ``ERROR: Checksum mismatch`` line otherwise) and prints the same two timing lines -- for a whole batch instead of one
vector (batch 1 reproduces run.c's setting).

    python -m structurednets_b200.speed_comparison.run [layer_description.txt] [--batch 1 256 65536] [--iterations 100]

Output, one block per batch size:
    dense_time: <ms per forward>ms
    sss_time: <ms per forward>ms
"""
import argparse
import sys

import numpy as np
import torch

from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import SyntheticMixedSystem, _Stage

_SCALARS = ("input_size", "output_size", "nb_states", "max_state_space_dim", "max_input_dim", "max_output_dim")


def parse_layer_description(path: str) -> dict:
    """Parser for the text format of prepare_run.py:10-31 (``name rows cols`` + rows of comma-separated values; ``name len`` + one
    line for vectors; ``name value`` for the scalars of :54-55,70-73)."""
    out = {}
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        tok = lines[i].split()
        i += 1
        if not tok:
            continue
        name = tok[0]
        if name in _SCALARS:
            out[name] = int(tok[1])
        elif len(tok) == 2:                                   # vector
            out[name] = np.array([float(v) for v in lines[i].split(",") if v.strip()], dtype=np.float32)
            assert len(out[name]) == int(tok[1]), "vector %s: expected %s values" % (name, tok[1])
            i += 1
        elif len(tok) == 3:                                   # matrix
            r, c = int(tok[1]), int(tok[2])
            m = np.zeros((r, c), dtype=np.float32)
            for k in range(r):
                vals = [float(v) for v in lines[i + k].split(",") if v.strip()]
                assert len(vals) == c, "matrix %s row %d: expected %d values" % (name, k, c)
                m[k] = vals
            out[name] = m
            i += r
        else:
            raise ValueError("unexpected line in %s: %r" % (path, lines[i - 1][:80]))
    return out


def sss_layer_from_description(d: dict) -> SSSLayer:
    n = d["nb_states"]
    get = lambda name: [d["%s_%d" % (name, k)] for k in range(n)]
    A, B, C, D, E, F, G = (get(x) for x in "ABCDEFG")
    dims_in = [m.shape[1] for m in B]
    dims_out = [m.shape[0] for m in D]
    system = SyntheticMixedSystem(dims_in, dims_out, [_Stage(A[k].astype(np.float64), B[k].astype(np.float64), C[k].astype(np.float64), D[k].astype(np.float64)) for k in range(n)],
                                  [_Stage(E[k].astype(np.float64), F[k].astype(np.float64), G[k].astype(np.float64)) for k in range(n)])
    return SSSLayer(d["input_size"], d["output_size"], 1.0, initial_bias=d["sss_bias"].astype(np.float64), nb_states=n,
                    initial_system_approx=system)


def _check(y: np.ndarray, truth: np.ndarray, what: str) -> bool:
    err = np.abs(y - truth.reshape(1, -1))
    if float(err.max()) > 1e-3:                               # run.c:177,357
        r, c = np.unravel_index(int(err.argmax()), err.shape)
        print("ERROR: Checksum mismatch for the %s layer in dimension %d" % (what, c))
        print("%s checksum ground truth: %f" % (what, truth.reshape(-1)[c]))
        print("Computed Y value: %f" % y[r, c])
        print("Absolute Error: %f" % err[r, c])
        return False
    return True


def _time_ms(fn, iterations: int) -> float:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iterations):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iterations


def run(path: str = "layer_description.txt", batches=(1,), iterations: int = 100, device: str = "cuda") -> int:
    d = parse_layer_description(path)
    W = torch.tensor(d["W"], device=device)
    b = torch.tensor(d["standard_bias"], device=device)
    layer = sss_layer_from_description(d).to(device)
    u = torch.tensor(d["checksum_inp"].reshape(1, -1), device=device)
    for B in batches:
        X = u.repeat(int(B), 1).contiguous()                  # every row is the checksum vector: every output row is checked
        with torch.no_grad():
            dense_ms = _time_ms(lambda: torch.addmm(b, X, W.t()), iterations)
            sss_ms = _time_ms(lambda: layer(X), iterations)
            y_dense = torch.addmm(b, X, W.t()).cpu().numpy()
            y_sss = layer(X).cpu().numpy()
        print("batch %d" % B)
        print("dense_time: %fms" % dense_ms)
        print("sss_time: %fms " % sss_ms)
        if not _check(y_dense, d["standard_checksum_out"], "standard") or not _check(y_sss, d["sss_checksum_out"], "SSS"):
            return 1
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("path", nargs="?", default="layer_description.txt")
    ap.add_argument("--batch", type=int, nargs="+", default=[1])
    ap.add_argument("--iterations", type=int, default=100)
    a = ap.parse_args()
    sys.exit(run(a.path, a.batch, a.iterations))
