"""Micro-benchmark of the persistent 16-bit hi/lo GEMM (wf_gemm16.cu) at the benchmark shapes (CUDA events)."""
import sys, torch
sys.path.insert(0, ".")
from weatherforecast_stgcn_maml_b200 import _lib, synth
from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs
from oracle import ref_port as P

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

G, N, T = 15, 441, 24
R = N * T
rows = G * R
st = _lib.stream_ptr()
err = torch.zeros(1, dtype=torch.int32, device="cuda")
def w16(W, fmt=0):
    hi = torch.empty(W.shape, dtype=torch.int16, device="cuda"); lo = torch.empty_like(hi)
    _lib.call("wf_split16", _lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo), W.numel(), fmt, st)
    return hi, lo
import os
SHAPES = ((256, 256),) if os.environ.get("WF_G16_DEBUG") else ((256, 256), (256, 512), (128, 512), (512, 128))
for K, Nn in SHAPES:
    A = torch.randn(rows, K, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    hi, lo = w16(W); C = torch.empty(rows, Nn, device="cuda")
    t = timeit(lambda: _lib.call("wf_g16_gemm_nt", _lib.ptr(A), rows, 1, K, _lib.ptr(hi), _lib.ptr(lo), 0, Nn, None, None, 0, 0, 0, _lib.ptr(C), _lib.ptr(err), st))
    gb = (A.numel() + C.numel()) * 4 / 1e9
    print(f"rows-mode plain  K={K} N={Nn}: {t:7.1f} us  {gb / t * 1e6:7.0f} GB/s  {2 * rows * K * Nn * 3 / t / 1e6:7.0f} TF16/s-equiv")
# GCN layer with CSR
lats, lons = synth.region_grid(21, 21)
ei = P.knn_edges_canonical(lats, lons, 8)
graphs = StackedGraphs([RegionGraph(ei, R, "cuda") for _ in range(G)])
A = torch.randn(rows, 256, device="cuda").relu(); W = torch.randn(256, 256, device="cuda") / 16; b = torch.randn(256, device="cuda")
hi, lo = w16(W); Y = torch.empty(rows, 256, device="cuda")
rt = int(_lib.query("wf_transposed_pitch16", T, N))
YT = torch.zeros(G * 256 * rt, dtype=torch.int16, device="cuda"); YTl = torch.zeros_like(YT)
AGG = torch.empty(rows, 256, device="cuda")
for ct, pre in ((False, False), (False, True), (True, True)):
    t = timeit(lambda: _lib.call("wf_gcn_layer_fwd_g16", _lib.ptr(A), None, 0, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(b), _lib.ptr(graphs.rowptr), _lib.ptr(graphs.col), _lib.ptr(graphs.val), graphs.rowptr_stride, graphs.csr_stride, _lib.ptr(graphs.gather_rows) if pre else None, graphs.gather_max, graphs.gather_rows.shape[1], _lib.ptr(AGG), R, N, 256, 256, G, 1, 1, _lib.ptr(Y), _lib.ptr(YT) if ct else None, _lib.ptr(YTl) if ct else None, 0.0, None, 0, _lib.ptr(err), st))
    print(f"gcn layer (CSR, bias, relu, transposed={ct}, pre-aggregation={pre}): {t:7.1f} us")
print("err flag", int(err.item()))
