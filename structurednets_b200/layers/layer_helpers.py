"""Host-side helpers shared by the layers.

Mirrors ``structurednets/layers/layer_helpers.py`` of the reference: Glorot-uniform numpy
initialisation with limit sqrt(6 / sum(shape)) (layer_helpers.py:10-12) and the sparse-aware
parameter counter used by the reference's tests (layer_helpers.py:14-22).
"""
import numpy as np
import torch
import torch.nn as nn


def get_random_glorot_uniform_matrix(shape: tuple) -> np.ndarray:
    limit = np.sqrt(6 / sum(shape))
    return np.random.uniform(-limit, limit, size=shape)


def get_random_glorot_uniform_matrix_torch(shape: tuple) -> torch.Tensor:
    curr_mat = torch.tensor(get_random_glorot_uniform_matrix(shape=shape))
    curr_mat.requires_grad_()
    return curr_mat


def get_nb_model_parameters(model: nn.Module, count_gradientless_parameters=True) -> int:
    nb_parameters = 0
    for _name, param in model.named_parameters():
        if param.requires_grad or count_gradientless_parameters:
            nb_parameters += param._nnz() if param.is_sparse else param.numel()
    return nb_parameters
