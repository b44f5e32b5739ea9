"""Probe wf_tc_wgrad with one (a_k0, b_k0, klen, R) configuration per process."""
import sys
import torch
sys.path.insert(0, ".")
from weatherforecast_stgcn_maml_b200 import _lib
M, N, R, Bw, G = 512, 128, int(sys.argv[4]), 2, 2
a_k0, b_k0, klen = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
Z = G * Bw
AT = torch.randn(Z, M, R, device="cuda")
BT = torch.randn(Z, N, R, device="cuda")
BTlo = torch.empty_like(BT)
_lib.call("wf_split_lo", _lib.ptr(BT), _lib.ptr(BTlo), BT.numel(), _lib.stream_ptr())
dW = torch.zeros(G, M, N, device="cuda")
err = torch.zeros(1, dtype=torch.int32, device="cuda")
_lib.call("wf_tc_wgrad", _lib.ptr(AT), M, _lib.ptr(BT), _lib.ptr(BTlo), N, R, Bw, G, a_k0, b_k0, klen, _lib.ptr(dW),
          M * N, _lib.ptr(err), _lib.stream_ptr())
torch.cuda.synchronize()
A = AT[:, :, a_k0:a_k0 + klen].double().view(G, Bw, M, klen)
B = BT[:, :, b_k0:b_k0 + klen].double().view(G, Bw, N, klen) if b_k0 >= 0 else None
ref = torch.einsum("gwmk,gwnk->gmn", A, B)
print("cfg", sys.argv[1:], "err", int(err.item()), "rel", float((dW.double() - ref).abs().max() / ref.abs().max()))
