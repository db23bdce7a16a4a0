"""bf16 tcgen05 path of the low-rank layer against the float32 oracle -- needs a B200.
Stated bf16 tolerance: 2e-2 relative to the largest entry of the compared tensor (bf16 operands and bf16
hidden/outputs, fp32 accumulation)."""
import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200.layers.lr_layer import LRLayer

pytestmark = pytest.mark.gpu
TOL = 2e-2


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("shape", [(512, 2048, 1000, 0.1906), (300, 256, 72, 0.57), (8192, 2048, 1000, 0.1906)])
def test_lr_bf16_matches_fp32_oracle(shape, built_lib):
    B, i, o, share = shape
    np.random.seed(7)
    layer = LRLayer(i, o, share)
    rank = layer.left_lr.shape[1]
    assert rank % 8 == 0
    rng = np.random.default_rng(70 + B)
    X = rng.uniform(-1, 1, size=(B, i)).astype(np.float32)
    gy = rng.uniform(-1, 1, size=(B, o)).astype(np.float32) / B
    # oracle on the bf16-rounded features (what the kernel sees), fp32 parameters
    Xb = torch.tensor(X).bfloat16().float()
    gyb = torch.tensor(gy).bfloat16().float()
    L, R, b = [t.detach().clone().requires_grad_(True) for t in (layer.left_lr, layer.right_lr, layer.bias)]
    yo = O.lr_forward(Xb, L, R, b); (yo * gyb).sum().backward()
    layer = layer.to("cuda")
    y = layer(torch.tensor(X, device="cuda").bfloat16())
    assert y.dtype == torch.bfloat16 and y.shape == (B, o)
    (y.float() * torch.tensor(gy, device="cuda").bfloat16().float()).sum().backward()
    assert rel_err(y.detach().float().cpu().numpy(), yo.detach().numpy()) < TOL
    assert rel_err(layer.left_lr.grad.cpu().numpy(), L.grad.numpy()) < TOL
    assert rel_err(layer.right_lr.grad.cpu().numpy(), R.grad.numpy()) < TOL
    assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < TOL
    assert layer.left_lr.dtype == torch.float32 and layer.left_lr.grad.dtype == torch.float32   # fp32 master parameters


def test_lr_bf16_then_fp32_share_parameters(built_lib):
    np.random.seed(8)
    layer = LRLayer(256, 72, 0.57).to("cuda")
    x = torch.rand(64, 256, device="cuda") * 2 - 1
    y32 = layer(x)
    y16 = layer(x.bfloat16())
    assert rel_err(y16.detach().float().cpu().numpy(), y32.detach().cpu().numpy()) < TOL
    with torch.no_grad():
        layer.left_lr.mul_(2.0)          # bf16 copies must follow the fp32 master parameters
    y16b = layer(x.bfloat16())
    y32b = layer(x)
    assert rel_err(y16b.detach().float().cpu().numpy(), y32b.detach().cpu().numpy()) < TOL
