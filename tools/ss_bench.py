"""Timing of the SS-mode GEMMs at the benchmark shapes (15 tasks x 441 nodes x 24 steps), through the test entry points.
WF_SS_DEBUG=1 (no MMA) / 2 (no epilogue stores) / 3 ablate the pipeline."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weatherforecast_stgcn_maml_b200 import _lib  # noqa: E402

_lib.load()
G, Bw, T, N = 15, 1, 24, 441
tpw = (N + 127) // 128
blocks = G * Bw * T * tpw
err = torch.zeros(1, dtype=torch.int32, device="cuda")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def nodes(name, bn, avar, K, Ntot, fmt, kp=1):
    if avar == 1:
        a16 = torch.zeros(2, blocks * K * 128, dtype=torch.int16, device="cuda")
    else:
        a16 = torch.zeros(2, G * Bw * T * N * K, dtype=torch.int16, device="cuda")
    w = torch.zeros(2, G * Ntot * K, dtype=torch.int16, device="cuda")
    C = torch.empty(blocks * Ntot * 128, device="cuda")
    bias = torch.zeros(G, 2, Ntot, device="cuda") if os.environ.get("BIAS") else None
    b1, b2 = (_lib.ptr(bias[0, 0]), _lib.ptr(bias[0, 1])) if bias is not None else (None, None)
    us = timeit(lambda: _lib.call("wf_ss_nodes_gemm", bn, avar, _lib.ptr(a16), a16.shape[1], K, fmt, _lib.ptr(w[0]), _lib.ptr(w[1]),
                                  Ntot * K, Ntot, fmt, b1, b2, 2 * Ntot, _lib.ptr(C), T, N, Bw, G, kp, _lib.ptr(err), _lib.stream_ptr()))
    rows = G * Bw * T * N
    byts = rows * (K * 4 + Ntot * 4)
    print(f"{name:28s} bn={bn:3d} K={K:3d} N={Ntot:3d}: {us:7.1f} us  {byts / us / 1e3:7.0f} GB/s algorithmic")


nodes("P0 (feats -> gates)", 128, 0, 256, 512, 0)
nodes("P1 (h -> gates)", 256, 1, 128, 512, 0)
nodes("dX (dG -> dh)", 64, 1, 512, 128, 1)   # CTA pairs unless WF_SS_PAIRS=0 (then two 64-column parts on one CTA each)
nodes("dX, K split over 2 CTAs", 128, 1, 512, 128, 1, 2)
# weight gradients
dg = torch.zeros(2, blocks * 512 * 128, dtype=torch.int16, device="cuda")
h = torch.zeros(2, blocks * 128 * 128, dtype=torch.int16, device="cuda")
part = torch.empty(120 * 512 * 257, device="cuda")
buf = torch.empty(G, 2 * 512 * 128 + 512, device="cuda")
us = timeit(lambda: _lib.call("wf_ss_wgrad", _lib.ptr(dg), dg.shape[1], 2, _lib.ptr(h), h.shape[1], 0, 0, 0, 128, _lib.ptr(h), h.shape[1],
                              0, 1, 0, 128, T, N, Bw, G, _lib.ptr(part), part.numel(), _lib.ptr(buf), 128, 128,
                              _lib.ptr(buf[0, 65536:]), 128, 128, _lib.ptr(buf[0, 131072:]), buf.shape[1], _lib.ptr(err),
                              _lib.stream_ptr()))
print(f"{'wgrad [dW_ih | dW_hh]':28s}: {us:7.1f} us")
print("err", int(err.item()))
