"""Abstract base of every structured layer.

Keeps the reference's constructor contract (``structurednets/layers/structured_layer.py:11-37``):
the same argument asserts with the same messages, ``output_dim`` / ``use_bias`` attributes and a
float32 ``bias`` ``nn.Parameter`` of shape ``(output_dim,)`` registered first (so it is the first
``state_dict`` key), Glorot-uniform unless ``initial_bias`` is given.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch
import torch.nn as nn

from structurednets_b200.layers.layer_helpers import get_random_glorot_uniform_matrix


class StructuredLayer(nn.Module, ABC):
    def __init__(self, input_dim: int, output_dim: int, nb_params_share: float, use_bias=True,
                 initial_weight_matrix=None, initial_bias=None):
        super(StructuredLayer, self).__init__()
        assert input_dim > 0, "The input dim should be greater than 0"
        assert output_dim > 0, "The output dim should be greater than 0"
        if nb_params_share is not None:
            assert nb_params_share >= 0 and nb_params_share <= 1, "The nb_params_share should be between 0 and 1"

        self.output_dim = output_dim

        if initial_weight_matrix is not None:
            assert isinstance(initial_weight_matrix, np.ndarray), "The initial weight matrix must be passed as np.ndarray"
            assert len(initial_weight_matrix.shape) == 2, "The initial weight matrix should have 2 dimensions"
            assert np.array_equal(initial_weight_matrix.shape, np.array([output_dim, input_dim])), \
                "The initial weight matrix does not have the expected shape (" + str(output_dim) + "," + str(input_dim) + ")"

        self.use_bias = use_bias
        if use_bias:
            if initial_bias is not None:
                assert isinstance(initial_bias, np.ndarray), "The initial bias must be passed as np.ndarray"
                assert len(initial_bias.shape) == 1, "The initial bias is expected to be passed as vector"
                assert initial_bias.shape[0] == output_dim, \
                    "The initial bias has not the expected shape (it should contain " + str(output_dim) + " values in its first dimension)"
                bias = torch.tensor(initial_bias)
            else:
                bias = torch.tensor(get_random_glorot_uniform_matrix((output_dim,)))
            self.bias = nn.Parameter(bias.float())
        else:
            self.bias = None

    @abstractmethod
    def forward(self, U) -> torch.Tensor:
        return torch.zeros((U.shape[0], self.output_dim))

    @abstractmethod
    def get_nb_parameters(self) -> int:
        return 0

    # -- shared input checks for the CUDA paths (no CPU fallback by design) --------------------
    @staticmethod
    def _require_cuda(U: torch.Tensor, what: str):
        if not U.is_cuda:
            raise RuntimeError(
                what + ": this layer runs only on CUDA tensors (hand-written sm_100a kernels, no CPU fallback); "
                "move the module and the input to a B200 with .to('cuda')")
