"""SURVEY.md section 8f rank 2, measured: writes the reference's ``layer_description.txt`` for BASELINE config C1's layer (SSS
4096 -> 1000, 500 stages, statespace 16) with the oracle's output as checksum, runs the reference's own C program on it
(oracle/_ref/run: one vector on one host core, its own checksum test) and then the B200 counterpart
(structurednets_b200.speed_comparison.run) for batches 1 / 256 / 65536 with the same checksum protocol.

    python scripts/run_c_compare.py
"""
import os
import resource
import subprocess
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import layers_cpu as O  # noqa: E402
from oracle import run_c  # noqa: E402
from structurednets_b200.layers.sss_layer import SSSLayer  # noqa: E402
from structurednets_b200.speed_comparison.run import run  # noqa: E402
from structurednets_b200.synth import random_mixed_system  # noqa: E402


def main():
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=5000)
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm)
    rng = np.random.default_rng(1)
    u = rng.uniform(-1, 1, size=(1, 4096)).astype(np.float32)
    lists = [[p.detach() for p in getattr(layer, n)] for n in "ABCDEFG"]
    y = O.sss_forward(torch.tensor(u), *lists, layer.bias.detach(), layer.dims_in, layer.dims_out).numpy().reshape(-1)
    W = sysm.to_matrix().astype(np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "layer_description.txt")
        run_c.write_layer_description(path, *[[p.numpy() for p in l] for l in lists], layer.bias.detach().numpy(), u, y, W=W,
                                      standard_bias=layer.bias.detach().numpy())
        print("layer_description.txt: %.1f MB" % (os.path.getsize(path) / 1e6))
        binary = run_c.build_ref()
        if binary is not None:
            def unlimited_stack():   # run.c keeps W (16 MB) and the padded stage arrays on the stack
                resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
            r = subprocess.run([binary], cwd=tmp, capture_output=True, text=True, preexec_fn=unlimited_stack, timeout=1200)
            print("reference run.c (1 vector, 1 host core): rc=%d\n%s" % (r.returncode, r.stdout.strip()))
        if torch.cuda.is_available():
            print("B200 counterpart:")
            rc = run(path, batches=(1, 256, 65536), iterations=20)
            print("rc=%d" % rc)


if __name__ == "__main__":
    main()
