"""SURVEY.md section 8f rank 1, measured: the reference-shaped loop ``training_helpers.train`` (per-batch numpy slice -> tensor ->
H2D, per-batch ``.cpu()`` in the evaluation passes, one optimizer update per parameter tensor) against ``train_resident`` (features
uploaded once, device-side shuffle gather, loss / accuracy accumulated on the device) with and without the fused flat-buffer SGD,
on BASELINE config C1: SSSLayer 4096 -> 1000, 500 stages, statespace 16, batch 256, synthetic features held in host numpy.

    python scripts/train_loop_bench.py [--samples 8192] [--batch 256]

Prints one JSON line: wall seconds of whole calls with 2, 6 and 10 epochs (2 start evaluations + epochs, each epoch = training pass +
evaluation pass(es)), the steady-state seconds per epoch from the difference of the last two and the training samples per second it implies."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structurednets_b200 import training_helpers as TH  # noqa: E402
from structurednets_b200.layers.sss_layer import SSSLayer  # noqa: E402
from structurednets_b200.synth import random_mixed_system  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    rng = np.random.default_rng(1000)
    X = rng.uniform(-1, 1, size=(args.samples, 4096)).astype(np.float32)
    y = rng.integers(0, 1000, size=args.samples).astype(np.int64)
    Xv, yv = X[:1024], y[:1024]
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=5000)
    out = {}
    for name, fn, flat, graph, fused in (("train", TH.train, None, False, False), ("train_resident", TH.train_resident, None, False, False),
                                         ("train_resident_flat_sgd", TH.train_resident, TH.FlatSGD, False, False),
                                         ("train_resident_flat_sgd_cuda_graph", TH.train_resident, TH.FlatSGD, True, False),
                                         ("train_resident_flat_sgd_fused_loss_cuda_graph", TH.train_resident, TH.FlatSGD, True, True),
                                         ("train_adam", TH.train, "adam", False, False),
                                         ("train_resident_flat_adam_fused_loss_cuda_graph", TH.train_resident, TH.FlatAdam, True, True)):
        layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm).to("cuda")
        opt = torch.optim.Adam if flat == "adam" else (flat.for_model(layer) if flat is not None else torch.optim.SGD)
        kw = dict(X_val=Xv, y_val=yv, patience=1, batch_size=args.batch, lr=1e-3, restore_best_model=False, min_patience_improvement=1e6,
                  optimizer_class=opt, use_gpu=True)
        if graph:
            kw["cuda_graph"] = True
        if fused:
            kw["fused_loss"] = True
        np.random.seed(0)
        fn(layer, X[:2 * args.batch], y[:2 * args.batch], **kw)     # warm-up (plans, allocator)
        torch.cuda.synchronize()
        times = {}
        for pat in (1, 5, 9):        # 2, 6 and 10 epochs: the difference of the last two is 4 steady-state epochs (training + evaluation)
            kw["patience"] = pat
            np.random.seed(0)
            t0 = time.perf_counter()
            res = fn(layer, X, y, **kw)
            torch.cuda.synchronize()
            times[pat] = (time.perf_counter() - t0, len(res[5]))
        per_epoch = (times[9][0] - times[5][0]) / (times[9][1] - times[5][1])
        out[name] = dict(seconds_2_epochs=round(times[1][0], 4), seconds_6_epochs=round(times[5][0], 4), seconds_10_epochs=round(times[9][0], 4), seconds_per_epoch=round(per_epoch, 4),
                         train_samples_per_s=round(args.samples / per_epoch, 1), final_train_loss=float(res[5][-1]))
    out["config"] = dict(workload="C1: SSS 4096->1000, 500 stages, d=16, fp32, batch %d, %d host samples" % (args.batch, args.samples))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
