import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from structurednets_b200 import training_helpers as TH
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system
rng = np.random.default_rng(6)
X = rng.uniform(-1, 1, size=(600, 128)).astype(np.float32)
y = rng.integers(0, 24, size=600).astype(np.int64)
for name, fn, optf, extra in (("train+Adam", TH.train, lambda l: torch.optim.Adam, {}),
                              ("resident+Adam", TH.train_resident, lambda l: torch.optim.Adam, {}),
                              ("resident+FlatAdam", TH.train_resident, lambda l: TH.FlatAdam.for_model(l), {}),
                              ("resident+Adam+fused", TH.train_resident, lambda l: torch.optim.Adam, dict(fused_loss=True)),
                              ("resident+FlatAdam+fused", TH.train_resident, lambda l: TH.FlatAdam.for_model(l), dict(fused_loss=True)),
                              ("resident+FlatAdam+fused+graph", TH.train_resident, lambda l: TH.FlatAdam.for_model(l), dict(fused_loss=True, cuda_graph=True))):
    np.random.seed(12)
    layer = SSSLayer(128, 24, 0.9, nb_states=12, initial_system_approx=random_mixed_system(128, 24, 12, 16, seed=12)).to("cuda")
    res = fn(layer, X, y, X_val=X[:100], y_val=y[:100], patience=2, batch_size=200, lr=1e-3, restore_best_model=False,
             min_patience_improvement=1e6, optimizer_class=optf(layer), use_gpu=True, **extra)
    print(name, [float(v) for v in res[5]], float(res[1]))
