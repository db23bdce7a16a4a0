import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from structurednets_b200 import _lib
from tests.test_sss_tc_plan import TC_CASES, make_tc
case = TC_CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
layer, X = make_tc(case)
dev = torch.device("cuda")
B = X.shape[0]
gy = np.random.default_rng(7).uniform(-1, 1, size=(B, case["o"])).astype(np.float32)
layer = layer.to(dev)
L = _lib.lib()
tc = layer._tc_plan(dev)
ps = ctypes.byref(tc["struct"])
nc = tc["host"]["chunks"].shape[0]
flat = layer.flat_parameters()
_lib.check(L.sn_sss_tc_build(ps, _lib.ptr(flat), _lib.ptr(tc["coef"]), _lib.stream_ptr()), "build")
Xd = torch.tensor(X, device=dev)
y = torch.empty((B, case["o"]), device=dev)
rbuf = torch.zeros(int(L.sn_sss_tc_rbuf_floats(ps, B)), device=dev)
states = torch.zeros(int(L.sn_sss_tc_states_floats(ps, B)), device=dev)
_lib.check(L.sn_sss_tc_forward(ps, _lib.ptr(tc["coef"]), _lib.ptr(Xd), Xd.stride(0), _lib.ptr(y), y.stride(0), _lib.ptr(layer.bias), _lib.ptr(rbuf), _lib.ptr(states), B, _lib.stream_ptr()), "fwd")
gyd = torch.tensor(gy, device=dev)
ws = torch.zeros(int(L.sn_sss_tc_backward_workspace_floats(ps, B)), device=dev)
g = torch.zeros_like(flat); gb = torch.zeros(case["o"], device=dev)
_lib.check(L.sn_sss_tc_backward(ps, _lib.ptr(flat), _lib.ptr(tc["coef"]), _lib.ptr(Xd), Xd.stride(0), _lib.ptr(gyd), gyd.stride(0), _lib.ptr(states), _lib.ptr(ws), _lib.ptr(g), _lib.ptr(gb), None, 0, B, _lib.stream_ptr()), "bwd")
torch.cuda.synchronize()
wsn = ws.cpu().numpy()
Lb = wsn[:nc*B*32].reshape(nc, B, 32)
dM = wsn[nc*B*32: nc*B*32 + nc*64*192].reshape(nc, 64, 192)
st = states.cpu().numpy().reshape(nc, B, 32)
print("L absmax", np.abs(Lb).max(), "states absmax", np.abs(st).max(), "gb", np.abs(gb.cpu().numpy()).max(), "g", float(g.abs().max()))
for j in range(nc):
    print("chunk", j, tc["host"]["chunks"][j])
    blocks = np.abs(dM[j]).reshape(4, 16, 6, 32).max(axis=(1, 3))
    print(np.array2string(blocks, precision=3))
    # reference dM in numpy
    ch = tc["host"]["chunks"][j]
    G = np.concatenate([np.pad(gy[:, ch[4]:ch[4]+32], ((0,0),(0, max(0, ch[4]+32-gy.shape[1])))) , Lb[j]], axis=1).astype(np.float64)
    Xp = np.pad(X, ((0,0),(0,192)))
    V = np.concatenate([Xp[:, ch[2]:ch[2]+32*ch[6]], st[j]], axis=1).astype(np.float64)
    ref = G.T @ V
    print(" ref blocks"); print(np.array2string(np.abs(ref).reshape(4,16,-1,32).max(axis=(1,3)), precision=3))
    print(" max err vs numpy ref", np.abs(dM[j][:, :V.shape[1]] - ref).max())
