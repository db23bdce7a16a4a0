"""Low-rank layer on B200.  Drop-in for ``structurednets.layers.lr_layer.LRLayer`` (reference
layers/lr_layer.py:10-52): parameters ``bias, left_lr (out x r), right_lr (r x in)`` with
``r = int(int(share*in*out) / (in+out))``; ``forward`` computes ``left_lr (right_lr U^T)`` + bias.

Two compute paths behind the same module, selected by the dtype of the input features:
  * float32 input  -> fp32 CUDA-core GEMMs (csrc/lr.cu), parity 1e-5 with the reference;
  * bfloat16 input -> bf16 tcgen05 tensor-core path with fp32 accumulation and fp32 master parameters
    (csrc/lr_tc.cu; BASELINE config C2), output bfloat16.
"""
import numpy as np
import torch
import torch.nn as nn

from structurednets_b200 import _lib
from structurednets_b200.layers.flat_params import FlatParamsMixin
from structurednets_b200.layers.layer_helpers import get_random_glorot_uniform_matrix
from structurednets_b200.layers.structured_layer import StructuredLayer


class _LRFunctionF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, anchor, layer):
        B = U.shape[0]
        out_dim, rank = layer.left_lr.shape
        in_dim = layer.right_lr.shape[1]
        hidden = torch.empty((B, max(rank, 1)), dtype=torch.float32, device=U.device)
        y = torch.empty((B, out_dim), dtype=torch.float32, device=U.device)
        rc = _lib.lib().sn_lr_forward_f32(_lib.ptr(U), U.stride(0), _lib.ptr(layer.left_lr), _lib.ptr(layer.right_lr),
                                          _lib.ptr(layer.bias if layer.use_bias else None), _lib.ptr(hidden), _lib.ptr(y),
                                          y.stride(0), B, in_dim, out_dim, rank, _lib.stream_ptr())
        _lib.check(rc, "sn_lr_forward_f32")
        ctx.layer = layer
        ctx.save_for_backward(U, hidden)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        layer = ctx.layer
        U, hidden = ctx.saved_tensors
        B = U.shape[0]
        out_dim, rank = layer.left_lr.shape
        in_dim = layer.right_lr.shape[1]
        grad_y = grad_y.contiguous().float()
        layer._prepare_grad_accumulation()
        gh = torch.empty_like(hidden)
        gx = torch.empty_like(U) if ctx.needs_input_grad[0] else None
        gl = layer.left_lr.grad if layer.left_lr.requires_grad else None
        gr = layer.right_lr.grad if layer.right_lr.requires_grad else None
        gb = layer.bias.grad if (layer.use_bias and layer.bias.requires_grad) else None
        rc = _lib.lib().sn_lr_backward_f32(_lib.ptr(U), U.stride(0), _lib.ptr(grad_y), grad_y.stride(0), _lib.ptr(layer.left_lr),
                                           _lib.ptr(layer.right_lr), _lib.ptr(hidden), _lib.ptr(gh), _lib.ptr(gl), _lib.ptr(gr),
                                           _lib.ptr(gb), _lib.ptr(gx), gx.stride(0) if gx is not None else 0, B, in_dim, out_dim,
                                           rank, _lib.stream_ptr())
        _lib.check(rc, "sn_lr_backward_f32")
        return gx, None, None


class LRLayer(FlatParamsMixin, StructuredLayer):
    def __init__(self, input_dim: int, output_dim: int, nb_params_share: float, use_bias=True, initial_weight_matrix=None,
                 initial_bias=None, initial_lr_components=None):
        super(LRLayer, self).__init__(input_dim=input_dim, output_dim=output_dim, nb_params_share=nb_params_share, use_bias=use_bias,
                                      initial_weight_matrix=initial_weight_matrix, initial_bias=initial_bias)

        assert initial_weight_matrix is None or initial_lr_components is None, "Either pass an initial weight matrix or initial lr components - not both"
        if initial_lr_components is not None:
            assert len(initial_lr_components) == 2, "Need to pass 2 lr components"
            assert initial_lr_components[0].shape[0] == output_dim and initial_lr_components[1].shape[1] == input_dim, "The provided lr components do not match the given input and output dimensions"
            assert initial_lr_components[0].shape[1] == initial_lr_components[1].shape[0], "The provided lr component shapes do not match - they can not be multiplied with each other"

        max_nb_parameters = int(nb_params_share * input_dim * output_dim)
        rank = int(max_nb_parameters / (input_dim + output_dim))

        if initial_lr_components is not None:
            left_lr = torch.tensor(np.asarray(initial_lr_components[0]))
            right_lr = torch.tensor(np.asarray(initial_lr_components[1]))
        elif initial_weight_matrix is not None:
            left, right = svd_low_rank_factors(initial_weight_matrix, rank)
            left_lr, right_lr = torch.tensor(left), torch.tensor(right)
        else:
            left_lr = torch.tensor(get_random_glorot_uniform_matrix((output_dim, rank)))
            right_lr = torch.tensor(get_random_glorot_uniform_matrix((rank, input_dim)))

        self.input_dim = input_dim
        self.left_lr = nn.Parameter(left_lr.float())
        self.right_lr = nn.Parameter(right_lr.float())
        self._flatten_parameters()

    def forward(self, U):
        self._require_cuda(U, "LRLayer.forward")
        assert U.dim() == 2 and U.shape[1] == self.right_lr.shape[1], "LRLayer expects a (batch, input_dim) input"
        self._ensure_flat()
        if U.stride(1) != 1:
            U = U.contiguous()
        anchor = self.__dict__.get("_dev_anchor")
        if anchor is None or anchor.device != U.device:
            anchor = torch.zeros(1, device=U.device, requires_grad=True)
            self.__dict__["_dev_anchor"] = anchor
        needs = torch.is_grad_enabled() and (self.left_lr.requires_grad or self.right_lr.requires_grad)
        if U.dtype == torch.bfloat16:
            from structurednets_b200.layers.lr_tc import lr_forward_bf16
            return lr_forward_bf16(self, U, anchor if needs else None)
        if U.dtype != torch.float32:
            U = U.float()
        return _LRFunctionF32.apply(U, anchor if needs else None, self)

    def get_nb_parameters(self) -> int:
        res = torch.numel(self.left_lr) + torch.numel(self.right_lr)
        if self.use_bias:
            res += torch.numel(self.bias)
        return int(res)


def svd_low_rank_factors(optim_mat: np.ndarray, rank: int):
    """Truncated-SVD factors with the singular values split as sqrt(S) on both sides, as the reference's
    LRApproximator does (approximators/lr_approximator.py:16-22)."""
    U, S, Vh = np.linalg.svd(optim_mat, full_matrices=False)
    S_root = np.sqrt(S)
    left = (U * S_root[None, :])[:, :rank]
    right = (S_root[:, None] * Vh)[:rank, :]
    return left, right
