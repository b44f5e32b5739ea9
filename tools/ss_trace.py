"""clock64 timeline of CTA 0 of wf_ss_kernel (build with NVCC_EXTRA=-DWF_SS_TRACE): per pipeline stage the cycle at which
the producer found the slot free, the MMA warp saw the data, and the MMA warp had issued the stage."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weatherforecast_stgcn_maml_b200 import _lib  # noqa: E402

lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "dx"
G, Bw, T, N = 15, 1, 24, 441
blocks = G * Bw * T * 4
err = torch.zeros(1, dtype=torch.int32, device="cuda")
if which == "wg":
    dg = torch.zeros(2, blocks * 512 * 128, dtype=torch.int16, device="cuda")
    h = torch.zeros(2, blocks * 128 * 128, dtype=torch.int16, device="cuda")
    part = torch.empty(120 * 512 * 257, device="cuda")
    buf = torch.empty(G, 2 * 512 * 128 + 512, device="cuda")
    for _ in range(3):
        _lib.call("wf_ss_wgrad", _lib.ptr(dg), dg.shape[1], 2, _lib.ptr(h), h.shape[1], 0, 0, 0, 128, _lib.ptr(h), h.shape[1], 0, 1, 0,
                  128, T, N, Bw, G, _lib.ptr(part), part.numel(), _lib.ptr(buf), 128, 128, _lib.ptr(buf[0, 65536:]), 128, 128,
                  _lib.ptr(buf[0, 131072:]), buf.shape[1], _lib.ptr(err), _lib.stream_ptr())
elif not which.startswith("gcn"):
    bn, avar, K, Ntot, fmt, kp = {"dx": (64, 1, 512, 128, 1, 1), "dx2": (128, 1, 512, 128, 1, 2), "p1": (256, 1, 128, 512, 0, 1),
                                  "p0": (128, 0, 256, 512, 0, 1)}[which]
    a16 = torch.zeros(2, blocks * K * 128 if avar else G * T * N * K, dtype=torch.int16, device="cuda")
    w = torch.zeros(2, G * Ntot * K, dtype=torch.int16, device="cuda")
    Cm = torch.empty(blocks * Ntot * 128, device="cuda")
    for _ in range(3):
        _lib.call("wf_ss_nodes_gemm", bn, avar, _lib.ptr(a16), a16.shape[1], K, fmt, _lib.ptr(w[0]), _lib.ptr(w[1]), Ntot * K, Ntot,
                  fmt, None, None, 0, _lib.ptr(Cm), T, N, Bw, G, kp, _lib.ptr(err), _lib.stream_ptr())
if which.startswith("gcn"):
    cin, cout, R = (24 if which == "gcn24" else 256), 256, T * N
    Z = G * Bw
    x = torch.randn(Z, R, cin, device="cuda")
    W = torch.randn(cout, cin, device="cuda")
    from weatherforecast_stgcn_maml_b200.functional import split_weight16
    hi, lo = split_weight16(W)
    b = torch.zeros(cout, device="cuda")
    y16 = torch.empty(2, Z * R * cout, dtype=torch.int16, device="cuda")
    xs16 = torch.empty(2, Z * R * cin, dtype=torch.int16, device="cuda")
    for _ in range(3):
        _lib.call("wf_gcn_layer_fwd_ss", _lib.ptr(x), None, None, _lib.ptr(xs16), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(b), None, None, None,
                  0, 0, 0, None, R, cin, cout, G, Bw, 1, _lib.ptr(y16), None, 0.0, None, 0, _lib.ptr(err), _lib.stream_ptr())
torch.cuda.synchronize()
buf = (C.c_longlong * 1536)()
lib.wf_ss_trace_read.argtypes = [C.c_void_p]
lib.wf_ss_trace_read(buf)
t0 = buf[0]
print(f"{which}: stage  slot_free  data_landed  issued   (cycles from the first event; deltas vs previous stage)")
prev = None
for i in range(48):
    row = [buf[e * 256 + i] - t0 for e in range(3)]
    extra = buf[3 * 256 + i] - buf[0 * 256 + i]
    d = "" if prev is None else "  d=" + "/".join(str(row[e] - prev[e]) for e in range(3))
    print(f"{i:3d} {row[0]:10d} {row[1]:10d} {row[2]:10d}   land-free={row[1] - row[0]:6d} issue={row[2] - row[1]:5d} load_issue={extra:5d}{d}")
    prev = row
print("epilogue (warp 2): tile  accumulator_complete  drained   (duration; gap to the next tile's accumulator)")
for i in range(12):
    e0, e1, e2 = buf[4 * 256 + i] - t0, buf[5 * 256 + i] - t0, buf[4 * 256 + i + 1] - t0
    print(f"{i:3d} {e0:10d} {e1:10d}   dur={e1 - e0:6d}  next-start={e2 - e1:6d}")
