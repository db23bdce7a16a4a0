"""Host logic of the training-loop mirror (reference training_helpers.py) with a plain CPU module, and the
reference's test_train_improvement on the CUDA layers (gpu)."""
import numpy as np
import pytest
import torch

from structurednets_b200 import training_helpers as TH


def toy(n=200, i=12, o=5, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(n, i)).astype(np.float32)
    W = rng.uniform(-1, 1, size=(i, o)).astype(np.float32)
    return X, (X @ W).argmax(1).astype(np.int64)


def test_get_batch_slices_like_reference():
    X, y = toy()
    xb, yb = TH.get_batch(X, y, batch_size=64, batch_i=3)
    assert xb.shape == (8, 12) and yb.shape == (8,) and xb.dtype == torch.float32 and yb.dtype == torch.int64
    np.testing.assert_array_equal(xb.numpy(), X[192:200])
    xf, yf = TH.get_full_batch(X, y)
    assert xf.shape == (200, 12)


def test_train_contract_nine_tuple_and_early_stopping():
    torch.manual_seed(0)
    X, y = toy()
    model = torch.nn.Linear(12, 5)
    res = TH.train(model, X, y, patience=2, batch_size=50, lr=5e-2, min_patience_improvement=1e6, optimizer_class=torch.optim.SGD)
    assert len(res) == 9
    trained, s_tl, s_ta, s_vl, s_va, tlh, tah, vlh, vah = res
    assert isinstance(trained, torch.nn.Linear) and trained is not model          # restore_best_model -> pickled snapshot
    assert len(tlh) == len(vlh) == 3                                              # patience 2 + huge min improvement -> 3 epochs
    np.testing.assert_allclose(tlh, vlh)                                          # reference quirk: "val" history is the train set (:151)
    assert tlh[-1] < s_tl
    res2 = TH.train(model, X, y, patience=1, batch_size=50, lr=1e-2, restore_best_model=False, min_patience_improvement=1e6)
    assert res2[0] is model


def test_train_with_decreasing_lr_returns_lists_of_ten():
    X, y = toy(n=80)
    res = TH.train_with_decreasing_lr(torch.nn.Linear(12, 5), X, y, patience=1, batch_size=80, min_patience_improvement=1e6)
    assert len(res) == 9 and all(len(h) == 10 for h in res[1:])


def test_one_hot_targets_accuracy_path():
    X, y = toy()
    Y = np.eye(5, dtype=np.float32)[y]
    loss, acc = TH.get_loss_and_accuracy_for_model(torch.nn.Linear(12, 5), X, Y, loss_function_class=torch.nn.MSELoss, batch_size=64)
    assert 0.0 <= float(acc) <= 1.0 and float(loss) > 0


@pytest.mark.gpu
def test_train_improvement_all_layers(built_lib):
    """reference tests/test_layers.py:76-104: 20x20, share 0.5, MSE to ones, SGD lr 1e-3, full batch, through train()."""
    from structurednets_b200.layers.hmat_layer import HMatLayer
    from structurednets_b200.layers.ldr_layer import LDRLayer
    from structurednets_b200.layers.lr_layer import LRLayer
    from structurednets_b200.layers.psm_layer import PSMLayer
    from structurednets_b200.layers.sss_layer import SSSLayer
    from structurednets_b200.layers.tl_layer import TLLayer
    from structurednets_b200.synth import random_mixed_system
    np.random.seed(0)
    n = 1000
    X = np.random.uniform(-1, 1, size=(n, 20)).astype(np.float32)
    Y = np.ones((n, 20), dtype=np.float32)
    Xt = torch.tensor(X, device="cuda")
    layers = [LRLayer(20, 20, 0.5), PSMLayer(20, 20, nb_params_share=0.5), HMatLayer(20, 20, 0.5), LDRLayer(20, 20, 0.5), TLLayer(20, 20, 0.5),
              SSSLayer(20, 20, 0.5, nb_states=10, initial_system_approx=random_mixed_system(20, 20, 10, 2, seed=3))]
    for layer in layers:
        layer = layer.to("cuda")
        before = float((layer(Xt).detach() - 1).abs().max())
        trained, *_ = TH.train(model=layer, X_train=X, y_train=Y, patience=10, batch_size=n, verbose=False, lr=1e-3, restore_best_model=False,
                               loss_function_class=torch.nn.MSELoss, min_patience_improvement=1e6, optimizer_class=torch.optim.SGD, use_gpu=True)
        after = float((trained(Xt).detach() - 1).abs().max())
        assert after < before, type(layer).__name__ + " failed to improve the error"
