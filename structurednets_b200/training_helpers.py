"""The call sites of the hot path: the reference's feature-cache training loop, kept signature-compatible.

Mirror of ``structurednets/training_helpers.py`` (reference :1-181): ``get_batch`` (numpy slice -> torch tensor ->
device, :31-40), ``get_loss_and_accuracy_for_model`` (:57-73), ``train`` (:107-181: Adam/SGD over
``model.parameters()``, early stopping on patience, best-model snapshot by ``pickle``, 9-tuple return) and
``train_with_decreasing_lr`` (:75-105: ten rounds, lr halved from 1).  The loop bodies call ``model(X)`` /
``loss.backward()`` / ``optimizer.step()`` exactly like the reference, so any of the layers in
``structurednets_b200.layers`` drops in.

Two reference quirks are kept because callers can observe them (SURVEY.md "smaller reference bugs"):
the per-epoch "validation" history is computed on the TRAINING set (:151), and ``use_gpu`` silently falls
back to the CPU device string when CUDA is missing (models/visionmodel.py:5-9) -- the layers themselves then
raise, because they have no CPU path.
"""
import math
import pickle

import numpy as np
import torch
from sklearn.model_selection import train_test_split
from sklearn.utils import shuffle


def get_device(use_gpu=True):
    """reference models/visionmodel.py:5-9"""
    if use_gpu:
        return torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    return 'cpu'


def load_features(path: str):
    """reference asset_helpers.py:41-46: pickle of [X (N,1,F) or (N,F), y, y_pred]"""
    with open(path, "rb") as f:
        X, y, y_pred = pickle.load(f)
    return np.squeeze(X), y, y_pred


def get_number_correct_predictions(y_true: torch.Tensor, y_pred: torch.Tensor):
    return torch.sum(y_pred.argmax(axis=1) == y_true)


def get_accuracy_from_nb_correct_predictions(nb_correct_predictions: torch.Tensor, nb_samples: int):
    return (nb_correct_predictions / nb_samples).cpu().detach().numpy()


def get_accuracy(y_true: torch.Tensor, y_pred: torch.Tensor):
    return get_accuracy_from_nb_correct_predictions(get_number_correct_predictions(y_true=y_true, y_pred=y_pred), nb_samples=len(y_true))


def get_loss(y_true: torch.Tensor, y_pred: torch.Tensor, loss_function_class=torch.nn.CrossEntropyLoss):
    return loss_function_class()(y_pred, target=y_true).cpu().detach().numpy()


def get_train_data(features_path: str):
    X, y, _ = load_features(features_path)
    X_train, X_val, y_train, y_val = train_test_split(X, y, test_size=0.2, random_state=42, shuffle=True)
    return X_train, X_val, y_train, y_val


def get_batch(X: np.ndarray, y: np.ndarray, batch_size: int, batch_i: int, use_gpu=False):
    lo = int(batch_i * batch_size)
    X_batch = X[lo:min(int((batch_i + 1) * batch_size), X.shape[0])]
    y_batch = y[lo:min(int((batch_i + 1) * batch_size), y.shape[0])]
    X_batch_t, y_batch_t = map(torch.tensor, (X_batch, y_batch))
    device = get_device(use_gpu=use_gpu)
    return X_batch_t.to(device), y_batch_t.to(device)


def get_full_batch(X: np.ndarray, y: np.ndarray, use_gpu=False):
    return get_batch(X=X, y=y, batch_size=X.shape[0], batch_i=0, use_gpu=use_gpu)


def transform_feature_dtypes(X_train, X_val, y_train, y_val):
    return X_train.astype(np.float32), X_val.astype(np.float32), y_train.astype(np.int64), y_val.astype(np.int64)


def train_with_features(model, features_path: str, patience=10, batch_size=1000, verbose=False, lr=1e-6, restore_best_model=True,
                        loss_function_class=torch.nn.CrossEntropyLoss, use_gpu=False):
    X_train, X_val, y_train, y_val = transform_feature_dtypes(*get_train_data(features_path))
    return train(model=model, X_train=X_train, X_val=X_val, y_train=y_train, y_val=y_val, patience=patience, batch_size=batch_size,
                 verbose=verbose, lr=lr, restore_best_model=restore_best_model, loss_function_class=loss_function_class, use_gpu=use_gpu)


def get_loss_and_accuracy_for_model(model, X: np.ndarray, y: np.ndarray, loss_function_class=torch.nn.CrossEntropyLoss, batch_size=1000,
                                    use_gpu=False):
    nb_batches_per_epoch = np.ceil(X.shape[0] / batch_size).astype("int")
    cumulated_loss = 0
    nb_correct_predictions = 0
    with torch.no_grad():   # evaluation only (the reference builds and drops a graph here)
        for batch_i in range(nb_batches_per_epoch):
            X_batch_t, y_batch_t = get_batch(X, y, batch_size=batch_size, batch_i=batch_i, use_gpu=use_gpu)
            y_pred = model(X_batch_t)
            cumulated_loss += get_loss(y_batch_t, y_pred, loss_function_class=loss_function_class)
            target = y_batch_t.argmax(axis=1) if (len(y.shape) == 2 and y.shape[1] > 1) else y_batch_t
            nb_correct_predictions += get_number_correct_predictions(y_true=target, y_pred=y_pred)
    loss = cumulated_loss / nb_batches_per_epoch
    accuracy = get_accuracy_from_nb_correct_predictions(nb_correct_predictions=nb_correct_predictions, nb_samples=len(y))
    return loss, accuracy


def _dp_world(grad_sync):
    """(group, world size) of a data-parallel run, or (None, 1).  Only a ``grad_sync`` with a process group behind it counts."""
    if grad_sync is None:
        return None, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 1
    group = getattr(grad_sync, "group", None)
    return group, dist.get_world_size(group)


def _dp_check_equal_steps(grad_sync, nb_batches_per_epoch, device):
    """Every rank must run the same number of optimizer steps per epoch: a rank that stops early leaves the others waiting in
    the all-reduce of the next step."""
    group, world = _dp_world(grad_sync)
    if world == 1:
        return
    import torch.distributed as dist
    t = torch.tensor([int(nb_batches_per_epoch), -int(nb_batches_per_epoch)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    assert int(t[0]) == -int(t[1]), ("data-parallel training needs the same number of batches per epoch on every rank "
                                     "(got between %d and %d): shard the samples evenly or pad the shards" % (-int(t[1]), int(t[0])))


def _dp_reduce_stats(grad_sync, loss, accuracy, nb_samples, device):
    """Loss and accuracy of the GLOBAL data set from every rank's shard values, identical on all ranks, so that early stopping and
    the best-model choice (reference :161-170) are taken in lockstep.  Loss: mean of the ranks' per-shard values (each is a mean over
    equally many batches); accuracy: correct predictions over all samples."""
    group, world = _dp_world(grad_sync)
    if world == 1:
        return loss, accuracy
    import torch.distributed as dist
    t = torch.tensor([float(loss), float(accuracy) * nb_samples, float(nb_samples)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    return np.float32(t[0] / world), np.float32(t[1] / t[2])


def train_with_decreasing_lr(model, X_train, y_train, X_val=None, y_val=None, patience=10, batch_size=1000, verbose=False,
                             loss_function_class=torch.nn.CrossEntropyLoss, min_patience_improvement=1e-10,
                             optimizer_class=torch.optim.SGD, use_gpu=False):
    if X_val is None or y_val is None:
        X_train, X_val, y_train, y_val = train_test_split(X_train, y_train, test_size=0.2)
    lr = 1   # halved after every round; the best model is restored between rounds (reference :81-103)
    trained_model = model
    histories = [[] for _ in range(8)]
    for _ in range(10):
        res = train(model=trained_model, X_train=X_train, y_train=y_train, X_val=X_val, y_val=y_val, patience=patience, batch_size=batch_size,
                    verbose=verbose, lr=lr, restore_best_model=True, loss_function_class=loss_function_class,
                    min_patience_improvement=min_patience_improvement, optimizer_class=optimizer_class, use_gpu=use_gpu)
        trained_model = res[0]
        for h, v in zip(histories, res[1:]):
            h.append(v)
        lr *= 0.5
    return (trained_model, *histories)


def train(model, X_train, y_train, X_val=None, y_val=None, patience=10, batch_size=1000, verbose=False, lr=1e-6, restore_best_model=True,
          loss_function_class=torch.nn.CrossEntropyLoss, min_patience_improvement=1e-10, optimizer_class=torch.optim.Adam, use_gpu=False,
          grad_sync=None):
    """Same contract as the reference ``train`` (:107-181).  ``grad_sync`` (optional, new): a callable invoked between
    ``loss.backward()`` and ``optimizer.step()`` -- ``structurednets_b200.distributed.GradSynchronizer`` for data-parallel
    training (one NCCL all-reduce of the flat gradient buffer).  With a process group behind ``grad_sync`` the epoch losses and
    accuracies are reduced over the ranks before the patience and best-model decisions, so all ranks stop in the same epoch and
    return the same model; ranks must hold equally many batches per epoch (asserted)."""
    if X_val is None or y_val is None:
        X_train, X_val, y_train, y_val = train_test_split(X_train, y_train, test_size=0.2)

    optimizer = optimizer_class(model.parameters(), lr=lr)
    loss_function = loss_function_class()
    nb_batches_per_epoch = np.ceil(X_train.shape[0] / batch_size).astype("int")
    stats_device = get_device(use_gpu=use_gpu)
    _dp_check_equal_steps(grad_sync, nb_batches_per_epoch, stats_device)
    evaluate = lambda X, y: _dp_reduce_stats(grad_sync, *get_loss_and_accuracy_for_model(
        model=model, X=X, y=y, loss_function_class=loss_function_class, batch_size=batch_size, use_gpu=use_gpu), len(y), stats_device)
    start_train_loss, start_train_accuracy = evaluate(X_train, y_train)
    start_val_loss, start_val_accuracy = evaluate(X_val, y_val)
    if verbose:
        print("------ Training Start -------")
        print("Start Train Loss: " + str(start_train_loss))
        print("Start Val Loss: " + str(start_val_loss))
        print("Start Train Accuracy: " + str(start_train_accuracy))
        print("Start Val Accuracy: " + str(start_val_accuracy))

    train_loss_history, train_accuracy_history, val_loss_history, val_accuracy_history = [], [], [], []
    best_val_loss = start_val_loss
    best_model = pickle.loads(pickle.dumps(model))   # the modules are picklable by contract

    continue_training = True
    epoch = 1
    while continue_training:
        X_train_shuffled, y_train_shuffled = shuffle(X_train, y_train)
        for batch_i in range(nb_batches_per_epoch):
            X_batch_t, y_batch_t = get_batch(X_train_shuffled, y_train_shuffled, batch_size=batch_size, batch_i=batch_i, use_gpu=use_gpu)
            outputs_train = model(X_batch_t)
            loss_train = loss_function(outputs_train, target=y_batch_t)
            optimizer.zero_grad()
            loss_train.backward()
            if grad_sync is not None:
                grad_sync()
            optimizer.step()

        train_loss, train_accuracy = evaluate(X_train, y_train)
        train_loss_history.append(train_loss)
        train_accuracy_history.append(train_accuracy)
        val_loss, val_accuracy = evaluate(X_train, y_train)   # sic: the reference evaluates the training set here (:151)
        val_loss_history.append(val_loss)
        val_accuracy_history.append(val_accuracy)
        if verbose:
            print("--- Epoch " + str(epoch) + " ---")
            print("Train Acc: " + str(train_accuracy_history[-1]))
            print("Train Loss: " + str(train_loss_history[-1]))
            print("Val Acc: " + str(val_accuracy_history[-1]))
            print("Val Loss: " + str(val_loss_history[-1]))

        if len(val_loss_history) > patience \
                and (np.min(val_loss_history[-patience:]) >= np.min(val_loss_history[:-patience]) - min_patience_improvement
                     or math.isnan(val_loss_history[-1])):
            continue_training = False
        if val_loss_history[-1] < best_val_loss:
            best_val_loss = val_loss_history[-1]
            best_model = pickle.loads(pickle.dumps(model))
            if verbose:
                print("Updated the best model found - new best val loss is " + str(best_val_loss))
        epoch += 1

    model_to_return = best_model if restore_best_model else model
    return (model_to_return, start_train_loss, start_train_accuracy, start_val_loss, start_val_accuracy, train_loss_history,
            train_accuracy_history, val_loss_history, val_accuracy_history)


# ----------------------------------------------------------------------------------------------------------------------
# Device-resident training loop.  Same signature, same 9-tuple, same early stopping and the same
# batch order as ``train`` for a given numpy seed -- but the features are uploaded ONCE (pinned, asynchronous), every epoch's
# shuffle is an index permutation gathered on the device, and loss / accuracy are accumulated in device tensors: one host
# synchronisation per evaluation pass instead of one per batch (reference training_helpers.py:17,24,31-40,147-153).
# ----------------------------------------------------------------------------------------------------------------------
def _to_device_once(a: np.ndarray, device):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if torch.device(device).type == "cuda":
        return t.pin_memory().to(device, non_blocking=True)
    return t.to(device)


def _evaluate_resident_fused(model, Xd: torch.Tensor, yd: torch.Tensor, fused: "FusedLoss", batch_size: int):
    """One device pass: every batch's loss and correct-prediction count are accumulated by the fused kernel into two device scalars,
    read once at the end (reference training_helpers.py:57-73 synchronises twice per batch)."""
    n = Xd.shape[0]
    nb_batches = np.ceil(n / batch_size).astype("int")
    fused.reset()
    with torch.no_grad():
        for batch_i in range(nb_batches):
            lo, hi = int(batch_i * batch_size), min(int((batch_i + 1) * batch_size), n)
            fused(model(Xd[lo:hi]), yd[lo:hi])
    both = torch.cat([fused.loss_sum, fused.correct.float()]).cpu().numpy()        # one synchronisation for the whole pass
    return np.float32(both[0]) / nb_batches, np.float32(both[1] / len(yd))


def _evaluate_resident(model, Xd: torch.Tensor, yd: torch.Tensor, loss_function, batch_size: int):
    """``get_loss_and_accuracy_for_model`` (:57-73) on resident tensors: the float32 batch losses are summed in the same order,
    the division by the (numpy int) batch count happens on the host like in the reference, so the returned values are identical."""
    n = Xd.shape[0]
    nb_batches = np.ceil(n / batch_size).astype("int")
    cumulated_loss, nb_correct = None, None
    one_hot = yd.dim() == 2 and yd.shape[1] > 1
    with torch.no_grad():
        for batch_i in range(nb_batches):
            lo, hi = int(batch_i * batch_size), min(int((batch_i + 1) * batch_size), n)
            y_pred = model(Xd[lo:hi])
            loss = loss_function(y_pred, target=yd[lo:hi]).detach()
            target = yd[lo:hi].argmax(axis=1) if one_hot else yd[lo:hi]
            correct = torch.sum(y_pred.argmax(axis=1) == target)
            cumulated_loss = loss if cumulated_loss is None else cumulated_loss + loss
            nb_correct = correct if nb_correct is None else nb_correct + correct
    loss = (0 + cumulated_loss.cpu().numpy()) / nb_batches          # one synchronisation for the whole pass
    accuracy = (nb_correct / len(yd)).cpu().detach().numpy()
    return loss, accuracy


class _FusedLossFn(torch.autograd.Function):
    """Loss value + gradient of the model output in ONE kernel (csrc/train.cu: sn_ce_loss / sn_mse_loss); the number of correct
    predictions of the batch is counted in the same pass (``counters``: device float loss sum, device int correct count -- both are
    accumulated into, so an evaluation pass reads them once at its end)."""

    @staticmethod
    def forward(ctx, outputs, target, kind, counters, want_grad):
        from structurednets_b200 import _lib
        L = _lib.lib()
        outputs = outputs.contiguous()
        B, C = outputs.shape
        loss_sum, correct = counters
        before = loss_sum.clone()
        grad = torch.empty_like(outputs) if want_grad else None
        if kind == "ce":
            rc = L.sn_ce_loss(_lib.ptr(outputs), outputs.stride(0), _lib.ptr(target), B, C, _lib.ptr(loss_sum), _lib.ptr(correct),
                              _lib.ptr(grad), grad.stride(0) if grad is not None else 0, _lib.stream_ptr())
        else:
            target = target.contiguous()
            rc = L.sn_mse_loss(_lib.ptr(outputs), outputs.stride(0), _lib.ptr(target), target.stride(0), B, C, _lib.ptr(loss_sum),
                               _lib.ptr(correct if C > 1 else None), _lib.ptr(grad), grad.stride(0) if grad is not None else 0, _lib.stream_ptr())
        _lib.check(rc, "fused loss")
        ctx.grad = grad
        return (loss_sum - before).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        return ctx.grad * grad_out, None, None, None, None


class FusedLoss:
    """Drop-in for ``torch.nn.CrossEntropyLoss()`` / ``torch.nn.MSELoss()`` (mean reduction, as the reference constructs them,
    training_helpers.py:110) on float32 CUDA outputs; also counts correct predictions (``.correct``) and sums batch losses
    (``.loss_sum``) on the device.  Anything else (CPU tensors, other dtypes) goes to the torch loss it wraps."""

    def __init__(self, loss_function_class, device):
        self.torch_loss = loss_function_class()
        self.kind = "ce" if isinstance(self.torch_loss, torch.nn.CrossEntropyLoss) else ("mse" if isinstance(self.torch_loss, torch.nn.MSELoss) else None)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=device)
        self.correct = torch.zeros(1, dtype=torch.int32, device=device)

    def usable(self, outputs, target):
        if self.kind is None or not outputs.is_cuda or outputs.dtype != torch.float32 or outputs.dim() != 2:
            return False
        if self.kind == "ce":
            return target.dtype == torch.int64 and target.dim() == 1
        return target.dtype == torch.float32 and target.shape == outputs.shape

    def reset(self):
        self.loss_sum.zero_()
        self.correct.zero_()

    def __call__(self, outputs, target):
        if not self.usable(outputs, target):
            return self.torch_loss(outputs, target=target)
        return _FusedLossFn.apply(outputs, target, self.kind, (self.loss_sum, self.correct), torch.is_grad_enabled() and outputs.requires_grad)


class FlatSGD:
    """Plain SGD over a layer's flat parameter / gradient buffers: one fused kernel (sn_flat_sgd) instead of one update per parameter
    tensor (the SSS layer has 3 501).  Loose parameters (sparse / float64 ones that do not live in the flat buffer) keep torch's SGD,
    so the reference's "sparse parameters: SGD only" rule (psm_layer.py:14-15) is unchanged.  ``grad_scale`` folds the data-parallel
    1 / world-size factor into the update (use ``GradSynchronizer(model)`` without ``scale`` then).  Pass as
    ``optimizer_class=FlatSGD.for_model(model)``."""

    def __init__(self, model, params, lr, grad_scale=1.0):
        self.lr, self.grad_scale = lr, grad_scale
        self.flat_p = model.flat_parameters() if hasattr(model, "flat_parameters") else None
        self.flat_g = model.flat_grad() if hasattr(model, "flat_grad") else None
        self.model = model
        flat_ptr = (self.flat_p.data_ptr(), self.flat_p.data_ptr() + self.flat_p.numel() * 4) if self.flat_p is not None else (0, 0)
        loose = [p for p in params if p.is_sparse or (p.numel() > 0 and not (flat_ptr[0] <= p.data_ptr() < flat_ptr[1]))]   # empty boundary blocks: nothing to update
        self.loose_params = loose
        self.loose = self._loose_optimizer(loose, lr) if loose else None

    @staticmethod
    def _loose_optimizer(loose, lr):
        return torch.optim.SGD(loose, lr=lr)

    @classmethod
    def for_model(cls, model, grad_scale=1.0):
        return lambda params, lr: cls(model, list(params), lr, grad_scale=grad_scale)

    def zero_grad(self, set_to_none=True):
        if self.flat_g is not None:
            self.model.zero_flat_grad()
        if self.loose is not None:
            self.loose.zero_grad(set_to_none=set_to_none)

    def _flat_step(self):
        if self.flat_p.is_cuda:
            from structurednets_b200 import _lib
            _lib.check(_lib.lib().sn_flat_sgd(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), self.flat_p.numel(), float(self.lr), float(self.grad_scale),
                                              _lib.stream_ptr()), "sn_flat_sgd")
            torch.autograd.graph.increment_version(self.flat_p)      # the kernel wrote the parameters behind autograd's back
        else:
            with torch.no_grad():
                self.flat_p.add_(self.flat_g, alpha=-self.lr * self.grad_scale)

    def step(self):
        if self.flat_p is not None:
            self._flat_step()
        if self.loose is not None:
            if self.grad_scale != 1.0:
                with torch.no_grad():
                    for p in self.loose_params:
                        if p.grad is not None:
                            (p.grad._values() if p.grad.is_sparse else p.grad).mul_(self.grad_scale)
            self.loose.step()


class FlatAdam(FlatSGD):
    """Adam (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad -- what the reference's
    ``optimizer_class=torch.optim.Adam`` default of ``train`` uses, training_helpers.py:107,111) over the flat buffers in one kernel
    (sn_flat_adam); the step counter lives on the device so that the step can be captured in a CUDA graph.  Loose dense parameters keep
    torch's Adam; sparse ones raise like torch does (reference: "SGD only", psm_layer.py:14-15)."""

    def __init__(self, model, params, lr, grad_scale=1.0, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(model, params, lr, grad_scale=grad_scale)
        self.betas, self.eps = betas, eps
        if self.flat_p is not None:
            self.exp_avg = torch.zeros_like(self.flat_p)
            self.exp_avg_sq = torch.zeros_like(self.flat_p)
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.flat_p.device)
            self.step_host = 0

    @staticmethod
    def _loose_optimizer(loose, lr):
        return torch.optim.Adam(loose, lr=lr)

    # the warm-up passes of a CUDA-graph capture must leave no trace in the moment estimates or the step counter
    def snapshot_state(self):
        return None if self.flat_p is None else (self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone(), self.step_host)

    def restore_state(self, snap):
        if snap is not None:
            with torch.no_grad():
                self.exp_avg.copy_(snap[0]); self.exp_avg_sq.copy_(snap[1]); self.step_dev.copy_(snap[2])
            self.step_host = snap[3]

    def _flat_step(self):
        if self.flat_p.is_cuda:
            from structurednets_b200 import _lib
            _lib.check(_lib.lib().sn_flat_adam(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                               self.flat_p.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                               float(self.grad_scale), _lib.ptr(self.step_dev), 0, _lib.stream_ptr()), "sn_flat_adam")
            torch.autograd.graph.increment_version(self.flat_p)      # the kernel wrote the parameters behind autograd's back
        else:
            self.step_host += 1
            with torch.no_grad():
                g = self.flat_g * self.grad_scale
                self.exp_avg.mul_(self.betas[0]).add_(g, alpha=1 - self.betas[0])
                self.exp_avg_sq.mul_(self.betas[1]).addcmul_(g, g, value=1 - self.betas[1])
                bc1, bc2 = 1 - self.betas[0] ** self.step_host, 1 - self.betas[1] ** self.step_host
                self.flat_p.addcdiv_(self.exp_avg, (self.exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(self.eps), value=-self.lr / bc1)


class _GraphedStep:
    """One training step (gather the batch by index, forward, loss, backward, optimizer update) captured once in a CUDA graph and
    replayed per batch: at small batches the step is a few hundred microseconds of kernels inside a millisecond of Python / launch
    overhead.  Needs an optimizer whose update is capturable and allocation-free (``FlatSGD``) and a fixed batch size; the ragged
    last batch of an epoch runs eagerly."""

    def __init__(self, model, optimizer, loss_function, Xd, yd, batch_size, grad_sync):
        self.sel = torch.zeros(batch_size, dtype=torch.int64, device=Xd.device)
        self.loss = None
        self.graph = torch.cuda.CUDAGraph()

        def body():
            outputs = model(Xd.index_select(0, self.sel))
            loss = loss_function(outputs, target=yd.index_select(0, self.sel))
            optimizer.zero_grad()
            loss.backward()
            if grad_sync is not None:
                grad_sync()
            optimizer.step()
            return loss.detach()

        params = [p.detach().clone() for p in model.parameters()]      # the warm-up steps must not train
        opt_state = optimizer.snapshot_state() if hasattr(optimizer, "snapshot_state") else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                body()
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for p, q in zip(model.parameters(), params):
                p.copy_(q)
        if hasattr(optimizer, "restore_state"):
            optimizer.restore_state(opt_state)
        with torch.cuda.graph(self.graph):
            self.loss = body()
        with torch.no_grad():                                            # the capture itself does not run the kernels, but be explicit
            for p, q in zip(model.parameters(), params):
                p.copy_(q)
        if hasattr(optimizer, "restore_state"):
            optimizer.restore_state(opt_state)

    def __call__(self, sel):
        self.sel.copy_(sel, non_blocking=True)
        self.graph.replay()
        return self.loss


def train_resident(model, X_train, y_train, X_val=None, y_val=None, patience=10, batch_size=1000, verbose=False, lr=1e-6,
                   restore_best_model=True, loss_function_class=torch.nn.CrossEntropyLoss, min_patience_improvement=1e-10,
                   optimizer_class=torch.optim.Adam, use_gpu=False, grad_sync=None, cuda_graph=False, fused_loss=False):
    """Drop-in for ``train`` with the data resident on the device.  Consumes the numpy RNG exactly like ``train`` (one
    ``sklearn.utils.shuffle`` per epoch), so both loops see the same batches for the same seed.  ``cuda_graph=True`` replays the
    full-size training step from a CUDA graph (see ``_GraphedStep``; use with ``FlatSGD`` / ``FlatAdam``).  ``fused_loss=True``
    computes loss, accuracy and the gradient of the model output in one kernel (``FusedLoss``; CrossEntropyLoss / MSELoss on CUDA):
    the same numbers to float32 rounding (the summation order differs from ATen's), not bit for bit."""
    if X_val is None or y_val is None:
        X_train, X_val, y_train, y_val = train_test_split(X_train, y_train, test_size=0.2)
    device = get_device(use_gpu=use_gpu)
    Xd, yd = _to_device_once(X_train, device), _to_device_once(y_train, device)
    Xv, yv = _to_device_once(X_val, device), _to_device_once(y_val, device)

    optimizer = optimizer_class(model.parameters(), lr=lr)
    fused = FusedLoss(loss_function_class, device) if (fused_loss and torch.device(device).type == "cuda") else None
    if fused is not None and fused.kind is None:
        fused = None
    loss_function = fused if fused is not None else loss_function_class()
    n = Xd.shape[0]
    nb_batches_per_epoch = np.ceil(n / batch_size).astype("int")
    _dp_check_equal_steps(grad_sync, nb_batches_per_epoch, device)
    if fused is not None:
        evaluate = lambda X, y: _dp_reduce_stats(grad_sync, *_evaluate_resident_fused(model, X, y, fused, batch_size), len(y), device)
    else:
        evaluate = lambda X, y: _dp_reduce_stats(grad_sync, *_evaluate_resident(model, X, y, loss_function, batch_size), len(y), device)
    start_train_loss, start_train_accuracy = evaluate(Xd, yd)
    start_val_loss, start_val_accuracy = evaluate(Xv, yv)
    if verbose:
        print("------ Training Start -------")
        print("Start Train Loss: " + str(start_train_loss))
        print("Start Val Loss: " + str(start_val_loss))

    train_loss_history, train_accuracy_history, val_loss_history, val_accuracy_history = [], [], [], []
    best_val_loss = start_val_loss
    best_model = pickle.loads(pickle.dumps(model))
    continue_training = True
    epoch = 1
    index = np.arange(n)
    graphed = None
    if cuda_graph and torch.device(device).type == "cuda" and n >= batch_size:
        graphed = _GraphedStep(model, optimizer, loss_function, Xd, yd, int(batch_size), grad_sync)
    while continue_training:
        perm = _to_device_once(shuffle(index), device)          # same RNG consumption as shuffle(X_train, y_train) in train()
        for batch_i in range(nb_batches_per_epoch):
            sel = perm[int(batch_i * batch_size):min(int((batch_i + 1) * batch_size), n)]
            if graphed is not None and len(sel) == batch_size:
                graphed(sel)
                continue
            outputs_train = model(Xd.index_select(0, sel))
            loss_train = loss_function(outputs_train, target=yd.index_select(0, sel))
            optimizer.zero_grad()
            loss_train.backward()
            if grad_sync is not None:
                grad_sync()
            optimizer.step()

        train_loss, train_accuracy = evaluate(Xd, yd)
        train_loss_history.append(train_loss)
        train_accuracy_history.append(train_accuracy)
        val_loss_history.append(train_loss)          # sic: the reference's "validation" pass runs on the training set (:151);
        val_accuracy_history.append(train_accuracy)  # it is the same deterministic forward pass, so it is not repeated
        if verbose:
            print("--- Epoch " + str(epoch) + " ---")
            print("Train Acc: " + str(train_accuracy_history[-1]))
            print("Train Loss: " + str(train_loss_history[-1]))

        if len(val_loss_history) > patience \
                and (np.min(val_loss_history[-patience:]) >= np.min(val_loss_history[:-patience]) - min_patience_improvement
                     or math.isnan(val_loss_history[-1])):
            continue_training = False
        if val_loss_history[-1] < best_val_loss:
            best_val_loss = val_loss_history[-1]
            best_model = pickle.loads(pickle.dumps(model))
        epoch += 1

    model_to_return = best_model if restore_best_model else model
    return (model_to_return, start_train_loss, start_train_accuracy, start_val_loss, start_val_accuracy, train_loss_history,
            train_accuracy_history, val_loss_history, val_accuracy_history)
