"""Tensor-core SSS path (csrc/sss_tc.cu) on a B200: every kernel's output through the C ABI against the torch emulator
of the chunked formulation, then the layer against the oracle at 1e-5 relative (north_star's fp32 tolerance)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import layers_cpu as O
from structurednets_b200 import _lib
from structurednets_b200.layers.sss_layer import SSSLayer
from structurednets_b200.synth import random_mixed_system
from tests import sss_tc_emulator as E
from tests.test_sss_tc_plan import TC_CASES, make_tc

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def lists32(layer):
    return [[p.detach().cpu().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]


@pytest.mark.parametrize("fused,chain", [(0, 1), (0, 0), (1, 1)])
@pytest.mark.parametrize("case", TC_CASES)
def test_each_kernel_against_the_emulator(case, fused, chain, built_lib, monkeypatch):
    """chain = 1: chunk scans on the tensor core (default), 0: SIMT scans; fused = 1: the scans-in-the-GEMM-epilogue forward."""
    monkeypatch.setenv("SNB200_SSS_TC_FUSED", str(fused))
    monkeypatch.setenv("SNB200_SSS_TC_CHAIN", str(chain))
    layer, X = make_tc(case)
    dev = torch.device("cuda")
    B = X.shape[0]
    rng = np.random.default_rng(7)
    gy = rng.uniform(-1, 1, size=(B, case["o"])).astype(np.float32)
    # ---- emulator (fp64 so it is the more accurate side) ----
    l64 = [[p.detach().double().clone().requires_grad_(True) for p in getattr(layer, n)] for n in "ABCDEFG"]
    chunks = E.make_chunks(layer.dims_in, layer.dims_out)
    mats = [tuple(m.detach().clone().requires_grad_(True) for m in E.chunk_matrices(*l64, layer.dims_in, layer.dims_out, k0, k1))
            for (k0, k1) in chunks]
    U = torch.tensor(X).double()
    b64 = layer.bias.detach().double()
    y_e, info = E.forward_chunked(U, *l64, b64, layer.dims_in, layer.dims_out, return_all=True, mats=mats)
    (y_e * torch.tensor(gy).double()).sum().backward()

    layer = layer.to(dev)
    L = _lib.lib()
    tc = layer._tc_plan(dev)
    ps = ctypes.byref(tc["struct"])
    nc = len(chunks)
    flat = layer.flat_parameters()
    # 1. build
    _lib.check(L.sn_sss_tc_build(ps, _lib.ptr(flat), _lib.ptr(tc["coef"]), _lib.stream_ptr()), "build")
    torch.cuda.synchronize()
    coef = tc["coef"].cpu().numpy()
    W = coef[:nc * 128 * 160].reshape(nc, 128, 160)
    SC = coef[nc * 128 * 160: nc * 128 * 160 + nc * 1536].reshape(nc, 1536)
    for j in range(nc):
        Wj = W[j, :64].astype(np.float64) + W[j, 64:].astype(np.float64)
        assert rel_err(Wj, mats[j][0].detach().numpy()) < 1e-6, f"W chunk {j}"
        assert np.all((W[j].view(np.uint32) & 0x1FFF) == 0), "W hi/lo parts must be tf32-representable"
        ref_sc = np.concatenate([mats[j][1].detach().numpy().reshape(-1), mats[j][2].detach().numpy().reshape(-1),
                                 mats[j][3].detach().numpy().reshape(-1), mats[j][4].detach().numpy().reshape(-1)])
        assert rel_err(SC[j], ref_sc) < 1e-6, f"SC chunk {j}"
    # 2./3. forward
    Xd = torch.tensor(X, device=dev)
    y = torch.empty((B, case["o"]), device=dev)
    nr = int(L.sn_sss_tc_rbuf_floats(ps, B))
    assert (nr == 0) == bool(fused)
    rbuf = torch.zeros(nr, device=dev) if nr else None
    states = torch.zeros(int(L.sn_sss_tc_states_floats(ps, B)), device=dev)
    _lib.check(L.sn_sss_tc_forward(ps, _lib.ptr(tc["coef"]), _lib.ptr(Xd), Xd.stride(0), _lib.ptr(y), y.stride(0), _lib.ptr(layer.bias),
                                   _lib.ptr(rbuf), _lib.ptr(states), B, _lib.stream_ptr()), "forward")
    torch.cuda.synchronize()
    st = states.cpu().numpy().reshape(nc, B, 32)
    if not fused:
        rb = rbuf.cpu().numpy()          # three dense arrays: Y [chunk][B][32], R [chunk][B][16], R' [chunk][B][16]
        NB = nc * B
        r = np.concatenate([rb[:NB * 32].reshape(nc, B, 32), rb[NB * 32:NB * 48].reshape(nc, B, 16), rb[NB * 48:NB * 64].reshape(nc, B, 16)], axis=2)
        lo = 32 if chain else 0      # the tensor-core scan overwrites yloc (columns 0..31) in place with yloc + O' e
        for j in range(nc):
            assert rel_err(r[j][:, lo:], info["loc"][j].detach().numpy()[:, lo:]) < RTOL, f"local GEMM chunk {j}"
    for j in range(nc):
        assert np.max(np.abs(st[j, :, :16] - info["s"][j].detach().numpy())) < 1e-5 * max(1.0, float(info["s"][j].abs().max())), f"s chunk {j}"
        assert np.max(np.abs(st[j, :, 16:] - info["e"][j + 1].detach().numpy())) < 1e-5 * max(1.0, float(info["e"][j + 1].abs().max())), f"e chunk {j}"
    assert rel_err(y.cpu().numpy(), y_e.detach().numpy()) < RTOL
    # 4./5./6. backward
    gyd = torch.tensor(gy, device=dev)
    ws = torch.zeros(int(L.sn_sss_tc_backward_workspace_floats(ps, B)), device=dev)
    g = torch.zeros_like(flat)
    gb = torch.zeros(case["o"], device=dev)
    _lib.check(L.sn_sss_tc_backward(ps, _lib.ptr(flat), _lib.ptr(tc["coef"]), _lib.ptr(Xd), Xd.stride(0), _lib.ptr(gyd), gyd.stride(0),
                                    _lib.ptr(states), _lib.ptr(ws), _lib.ptr(g), _lib.ptr(gb), None, 0, B, _lib.stream_ptr()), "backward")
    torch.cuda.synchronize()
    wsn = ws.cpu().numpy()
    dM = wsn[nc * B * 32: nc * B * 32 + nc * 64 * 192].reshape(nc, 64, 192)
    host = tc["host"]["chunks"]
    for j in range(nc):
        ncols, nkb = int(host[j, 3]), int(host[j, 6])
        gr = lambda t: t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape))
        gW = gr(mats[j][0])
        scale = max(np.max(np.abs(gW)), 1e-30)
        nrows = int(host[j, 5])
        valid = np.r_[0:nrows, 32:64]          # rows nrows..31 of dM hold the next chunk's grad_y columns: never read
        assert np.max(np.abs(dM[j][valid, :ncols] - gW[valid, :ncols])) / scale < RTOL, f"dM (W part) chunk {j}"
        c0 = nkb * 32
        assert np.max(np.abs(dM[j][32:48, c0:c0 + 16] - gr(mats[j][1]))) / scale < RTOL, f"dPhi chunk {j}"
        assert np.max(np.abs(dM[j][48:64, c0 + 16:c0 + 32] - gr(mats[j][2]))) / scale < RTOL, f"dPhi' chunk {j}"
        assert np.max(np.abs(dM[j][:nrows, c0:c0 + 16] - gr(mats[j][3])[:nrows])) / scale < RTOL, f"dO chunk {j}"
        assert np.max(np.abs(dM[j][:nrows, c0 + 16:c0 + 32] - gr(mats[j][4])[:nrows])) / scale < RTOL, f"dO' chunk {j}"
    assert rel_err(gb.cpu().numpy(), gy.sum(axis=0)) < RTOL
    # parameter gradients against the oracle
    l32 = lists32(layer)
    yo = O.sss_forward(torch.tensor(X), *l32, layer.bias.detach().cpu(), layer.dims_in, layer.dims_out)
    (yo * torch.tensor(gy)).sum().backward()
    offs = layer._param_offsets()
    gn = g.cpu().numpy()
    for li, name in enumerate("ABCDEFG"):
        got = np.concatenate([gn[offs[(name, k)]:offs[(name, k)] + p.numel()] for k, p in enumerate(l32[li])])
        ref = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)).reshape(-1) for p in l32[li]])
        if ref.size:
            assert rel_err(got, ref) < RTOL, f"grad {name}"


@pytest.mark.parametrize("B,fused", [(256, 0), (77, 1), (1000, 1), (1000, 0)])
def test_alexnet_last_layer_shape_tc_vs_oracle_and_simt(B, fused, built_lib, monkeypatch):
    """BASELINE config C1 (4096 -> 1000, 500 stages, statespace 16) through the module: tensor-core path vs oracle and vs the SIMT path."""
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=1001)
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm)
    assert layer._use_tc_path()
    rng = np.random.default_rng(1001)
    X = rng.uniform(-1, 1, size=(B, 4096)).astype(np.float32)
    gy = rng.uniform(-1, 1, size=(B, 1000)).astype(np.float32) / B
    dev = torch.device("cuda")
    l32 = lists32(layer)
    b = layer.bias.detach().clone().requires_grad_(True)
    yo = O.sss_forward(torch.tensor(X), *l32, b, layer.dims_in, layer.dims_out)
    (yo * torch.tensor(gy)).sum().backward()
    layer = layer.to(dev)
    Xd, gyd = torch.tensor(X, device=dev), torch.tensor(gy, device=dev)
    res = {}
    monkeypatch.setenv("SNB200_SSS_TC_FUSED", str(fused))
    for mode in ("tc", "simt"):
        monkeypatch.setenv("SNB200_SSS_PATH", mode)
        for p in layer.parameters():
            p.grad = None
        _lib.reset_launch_count()
        y = layer(Xd)
        (y * gyd).sum().backward()
        torch.cuda.synchronize()
        res[mode] = (y.detach().cpu().numpy(), layer.flat_grad().detach().cpu().numpy().copy(), _lib.launch_count())
        assert rel_err(res[mode][0], yo.detach().numpy()) < RTOL, mode
        for li, name in enumerate("ABCDEFG"):
            got = np.concatenate([p.grad.detach().cpu().numpy().reshape(-1) for p in getattr(layer, name)])
            ref = np.concatenate([(p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)).reshape(-1) for p in l32[li]])
            assert rel_err(got, ref) < RTOL, f"{mode}: grad {name}"
        assert rel_err(layer.bias.grad.cpu().numpy(), b.grad.numpy()) < RTOL
    assert 6 <= res["tc"][2] <= 9 and res["simt"][2] >= 3     # kernels launched: build, pack, gemm (+ scans) | scan, gemm, build-bwd
    assert rel_err(res["tc"][1], res["simt"][1]) < RTOL


@pytest.mark.parametrize("fused", [0, 1])
def test_tc_affine_map_at_full_batch(fused, built_lib, monkeypatch):
    """Size-independent property at B = 8192: SSS(a x1 + b x2) - bias = a (SSS(x1) - bias) + b (SSS(x2) - bias)."""
    monkeypatch.setenv("SNB200_SSS_PATH", "tc")
    monkeypatch.setenv("SNB200_SSS_TC_FUSED", str(fused))
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=5)
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm).to("cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    x1 = torch.rand((8192, 4096), device="cuda", generator=g) * 2 - 1
    x2 = torch.rand((8192, 4096), device="cuda", generator=g) * 2 - 1
    with torch.no_grad():
        y1, y2, y3 = layer(x1), layer(x2), layer(0.5 * x1 - 2.0 * x2)
        lin = 0.5 * (y1 - layer.bias) - 2.0 * (y2 - layer.bias) + layer.bias
    assert float((y3 - lin).abs().max() / lin.abs().max()) < 3e-5


def test_full_batch_gradients_tc_vs_simt(built_lib, monkeypatch):
    """BASELINE config C5's batch (65 536): the tensor-core path against the CUDA-core path (fp32 FMA accumulation), outputs and the
    whole flat gradient.  Guards the accumulation-length bound of the gradient GEMM (the tensor core adds into its fp32 accumulator
    with truncation; see csrc/sss_tc.cu) -- an error the small-batch oracle comparisons cannot see."""
    B = 65536
    sysm = random_mixed_system(4096, 1000, 500, 16, seed=5000)
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=sysm).to("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
    gy = (torch.rand((B, 1000), device="cuda", generator=g) * 2 - 1) / B
    res = {}
    # "tc": the default scans of the tensor-core path (warp-level tensor-core scans); "tc_chain": the tcgen05 chain kernels forced
    for mode in ("tc", "tc_chain", "simt"):
        monkeypatch.setenv("SNB200_SSS_PATH", mode.split("_")[0])
        if mode == "tc_chain":
            monkeypatch.setenv("SNB200_SSS_TC_CHAIN", "1")
        else:
            monkeypatch.delenv("SNB200_SSS_TC_CHAIN", raising=False)
        layer.zero_flat_grad()
        for p in layer.parameters():
            p.grad = None
        y = layer(x)
        y.backward(gy)
        torch.cuda.synchronize()
        res[mode] = (y.detach().double(), layer.flat_grad().detach().double().clone())
        del y
    for mode in ("tc", "tc_chain"):
        ey = float((res[mode][0] - res["simt"][0]).abs().max() / res["simt"][0].abs().max())
        eg = float((res[mode][1] - res["simt"][1]).abs().max() / res["simt"][1].abs().max())
        assert ey < RTOL and eg < RTOL, (mode, ey, eg)


@pytest.mark.parametrize("chain", ["0", "1"])
def test_forward_states_and_outputs_are_bitwise_reproducible(built_lib, monkeypatch, chain):
    """The forward has no atomics: y and the saved states S[chunk][B][32] must be bit-identical from run to run.  Guards the buffer
    hand-offs of the chain kernels (an mbarrier.arrive that released a ring slot to the next TMA load before the ld.shared of the
    writer warps had returned corrupted a few rows of S about once in eight passes at this size; scripts/stress_states.py)."""
    monkeypatch.setenv("SNB200_SSS_PATH", "tc")
    monkeypatch.setenv("SNB200_SSS_TC_CHAIN", chain)        # 1: tcgen05 chain kernels, 0: warp-level tensor-core scans (the default)
    B = 65536
    layer = SSSLayer(4096, 1000, 0.105, nb_states=500, initial_system_approx=random_mixed_system(4096, 1000, 500, 16, seed=5000)).to("cuda")
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((B, 4096), device="cuda", generator=g) * 2 - 1
    L = _lib.lib()
    layer._ensure_flat()
    tc = layer._tc_plan(dev)
    ps = ctypes.byref(tc["struct"])
    _lib.check(L.sn_sss_tc_build(ps, _lib.ptr(layer.flat_parameters()), _lib.ptr(tc["coef"]), _lib.stream_ptr()), "build")
    ref = None
    for rep in range(10):
        y = torch.empty((B, 1000), device=dev)
        rbuf = torch.full((int(L.sn_sss_tc_rbuf_floats(ps, B)),), float("nan"), device=dev)
        states = torch.full((int(L.sn_sss_tc_states_floats(ps, B)),), float("nan"), device=dev)
        _lib.check(L.sn_sss_tc_forward(ps, _lib.ptr(tc["coef"]), _lib.ptr(x), x.stride(0), _lib.ptr(y), y.stride(0), _lib.ptr(layer.bias),
                                       _lib.ptr(rbuf), _lib.ptr(states), B, _lib.stream_ptr()), "forward")
        torch.cuda.synchronize()
        assert bool(torch.isfinite(states).all()) and bool(torch.isfinite(y).all())
        if ref is None:
            ref = (y.clone(), states.clone())
        else:
            assert torch.equal(y, ref[0]), "y differs in repetition %d" % rep
            assert torch.equal(states, ref[1]), "saved states differ in repetition %d" % rep


@pytest.mark.parametrize("path", ["tc", "simt"])
def test_backward_after_a_parameter_update_raises(built_lib, monkeypatch, path):
    """The backward reads layer-wide buffers and the current parameters: an in-place parameter write between a forward and its backward
    must raise instead of returning gradients at the wrong point; the ordinary order (and gradient accumulation over two forwards) must not."""
    monkeypatch.setenv("SNB200_SSS_PATH", path)
    layer, X = make_tc(TC_CASES[2])
    layer = layer.to("cuda")
    x = torch.tensor(X, device="cuda")
    y1 = layer(x)
    y2 = layer(x)
    y1.sum().backward()
    y2.sum().backward()          # two forwards, two backwards, no update in between: fine
    y3 = layer(x)
    with torch.no_grad():
        layer.bias.add_(1.0)     # what an optimizer step does
    with pytest.raises(RuntimeError, match="modified in place"):
        y3.sum().backward()
    layer(x).sum().backward()    # a fresh forward works again
