"""tcgen05 3xTF32 GEMM (wf_tc_gemm_nt) against float64 matmul: FP32-class accuracy on the tensor cores."""
import pytest
import torch

from weatherforecast_stgcn_maml_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(rows_g, G, K, N, bias=False, relu=False, seed=0):
    torch.manual_seed(seed)
    A = torch.randn(G * rows_g, K, device="cuda")
    W = torch.randn(G, N, K, device="cuda") / K ** 0.5
    Wlo = torch.empty_like(W)
    b1 = torch.randn(G, N, device="cuda") if bias else None
    b2 = torch.randn(G, N, device="cuda") if bias else None
    C = torch.full((G * rows_g, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = _lib.stream_ptr()
    _lib.call("wf_split_lo", _lib.ptr(W), _lib.ptr(Wlo), W.numel(), st)
    _lib.call("wf_tc_gemm_nt", _lib.ptr(A), rows_g, G, K, _lib.ptr(W), _lib.ptr(Wlo), N * K, N, _lib.ptr(b1), _lib.ptr(b2),
              N, int(relu), _lib.ptr(C), _lib.ptr(err), st)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"pipeline timeout code {int(err.item())}"
    ref = torch.bmm(A.double().view(G, rows_g, K), W.double().transpose(1, 2))
    if bias:
        ref = ref + (b1 + b2).double()[:, None, :]
    if relu:
        ref = ref.clamp_min(0)
    ref = ref.view(G * rows_g, N)
    return float((C.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("rows_g,G,K,N", [(128, 1, 32, 128), (128, 1, 64, 256), (441, 1, 128, 512), (300, 3, 256, 256),
                                          (1000, 2, 128, 128), (10584, 1, 256, 512)])
def test_tc_gemm_matches_fp64(rows_g, G, K, N):
    assert _run(rows_g, G, K, N) <= 5e-6


def test_tc_gemm_bias_relu_epilogue():
    assert _run(441, 2, 256, 256, bias=True, relu=True) <= 5e-6
    assert _run(200, 1, 128, 512, bias=True) <= 5e-6


def _run16(rows_g, G, K, N, fmt, bias=False, relu=False, seed=0, scale=1.0):
    torch.manual_seed(seed)
    A = torch.randn(G * rows_g, K, device="cuda") * scale
    W = torch.randn(G, N, K, device="cuda") / K ** 0.5
    Whi = torch.empty(G, N, K, dtype=torch.int16, device="cuda")
    Wlo = torch.empty_like(Whi)
    b1 = torch.randn(G, N, device="cuda") * scale if bias else None
    b2 = torch.randn(G, N, device="cuda") * scale if bias else None
    C = torch.full((G * rows_g, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = _lib.stream_ptr()
    _lib.call("wf_split16", _lib.ptr(W), _lib.ptr(Whi), _lib.ptr(Wlo), W.numel(), fmt, st)
    _lib.call("wf_g16_gemm_nt", _lib.ptr(A), rows_g, G, K, _lib.ptr(Whi), _lib.ptr(Wlo), N * K, N, _lib.ptr(b1), _lib.ptr(b2),
              N, int(relu), fmt, _lib.ptr(C), _lib.ptr(err), st)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"pipeline timeout code {int(err.item())}"
    ref = torch.bmm(A.double().view(G, rows_g, K), W.double().transpose(1, 2))
    if bias:
        ref = ref + (b1 + b2).double()[:, None, :]
    if relu:
        ref = ref.clamp_min(0)
    ref = ref.view(G * rows_g, N)
    return float((C.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("rows_g,G,K,N", [(128, 1, 64, 128), (441, 1, 128, 512), (300, 3, 256, 256), (1000, 2, 128, 128),
                                          (10584, 2, 256, 512), (20000, 1, 256, 256), (70000, 1, 64, 128), (40000, 1, 128, 384)])
def test_g16_gemm_matches_fp64(rows_g, G, K, N):
    """Persistent 16-bit hi/lo GEMM: fp16 split ~2^-20, bf16 split ~2^-16 (many tiles per CTA at the larger sizes)."""
    assert _run16(rows_g, G, K, N, 0) <= 5e-6
    assert _run16(rows_g, G, K, N, 1) <= 5e-5
    assert _run16(rows_g, G, K, N, 1, scale=1e-7) <= 5e-5   # gradient-sized operands keep their precision in bf16


def test_g16_gemm_bias_relu_epilogue():
    assert _run16(441, 2, 256, 256, 0, bias=True, relu=True) <= 5e-6
    assert _run16(200, 1, 128, 512, 1, bias=True) <= 5e-5


def test_split16_reconstructs():
    x = torch.randn(4096, device="cuda") * 3.0
    # fp16: 22 significant bits down to the fp16 subnormal spacing (2^-24 absolute); bf16: 16 bits, fp32 range
    for fmt, dt, tol, floor in ((0, torch.float16, 2.0 ** -21, 2.0 ** -24), (1, torch.bfloat16, 2.0 ** -15, 0.0)):
        hi = torch.empty(4096, dtype=torch.int16, device="cuda")
        lo = torch.empty_like(hi)
        _lib.call("wf_split16", _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), fmt, _lib.stream_ptr())
        rec = hi.view(dt).float() + lo.view(dt).float()
        assert torch.equal(hi.view(dt), x.to(dt))
        assert bool(((rec - x).abs() <= (x.abs() * tol).clamp_min(floor)).all())


def test_split_lo_is_exact():
    x = torch.randn(4096, device="cuda") * 37.0
    lo = torch.empty_like(x)
    _lib.call("wf_split_lo", _lib.ptr(x), _lib.ptr(lo), x.numel(), _lib.stream_ptr())
    hi = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert torch.equal(hi + lo, x) and torch.all(lo.abs() <= x.abs() * 2.0 ** -10)


@pytest.mark.parametrize("nlat,nlon,T,G,Bw", [(5, 7, 6, 2, 2), (21, 21, 24, 1, 1), (12, 13, 5, 3, 1),
                                              (3, 43, 1, 1, 2), (2, 64, 3, 2, 1), (10, 13, 2, 1, 3)])
def test_engine_tensor_core_path_matches_fp32_path(nlat, nlon, T, G, Bw):
    """Whole window pass (GCN -> LSTM -> head -> MSE -> BPTT) on the tcgen05 path vs the exact FP32 SIMT path
    and vs the CPU oracle, v5 layer widths, ragged tiles (N not a multiple of 128, N = 128 exactly, a 1-row tile),
    single-step and two-step windows (no / one recurrent product), several tasks/windows."""
    from oracle import ref_port as P
    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.engine import (HybridEngine, V5Dims, flatten_trainable,
                                                        gcn_weights_from_state_dict, unflatten_trainable)
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs

    n, H = nlat * nlon, 3
    dims = V5Dims(num_nodes=n, window=T, horizon=H)
    lats, lons = synth.region_grid(nlat, nlon)
    eis = [P.knn_edges_ckdtree(lats, lons, 4) if g % 2 == 0 else P.knn_edges_canonical(lats, lons, 4) for g in range(G)]
    base = synth.init_v5_state_dict(9, gcn_bias_scale=0.05, horizon=H)
    sds = [{k: (v + 0.02 * torch.randn_like(v) * (g > 0) if k.startswith(("lstm.", "output_layer.")) else v)
            for k, v in base.items()} for g in range(G)]
    time_rows = T + H + 1 + Bw + 2
    feats = torch.stack([synth.synth_features(time_rows, n, 200 + g) for g in range(G)])
    per, per_task = n * 24, time_rows * n * 24
    starts = [[(g + 2 * b) % (Bw + 2) for b in range(Bw)] for g in range(G)]
    xo = torch.tensor([g * per_task + s * per for g in range(G) for s in starts[g]], device="cuda")
    to = xo + (T + 1) * per
    theta = torch.stack([flatten_trainable(sd, dims) for sd in sds]).cuda()
    graphs = StackedGraphs([RegionGraph(ei, dims.R, "cuda") for ei in eis])
    fd = feats.cuda()
    out = {}
    for prec in ("fp32", "tf32x3", "stepwise"):
        eng = HybridEngine(dims, G, Bw, "cuda", precision="fp32" if prec == "fp32" else "tf32x3",
                           lstm_mode="stepwise" if prec == "stepwise" else "persistent")
        assert eng.tc == (prec != "fp32") and eng.seq == (prec == "tf32x3")
        eng.gcn_forward(fd, 24, 0, xo, gcn_weights_from_state_dict(base, "cuda"), graphs)
        eng.check()
        eng.lstm_head_forward(theta, eng.P)
        eng.check()
        loss = eng.mse(feat=fd, tgt_off=to, feat_ld=24, grad_scale=1.0)
        hid = eng.hidden_states()
        grads = eng.backward(theta, eng.P)
        eng.check()
        out[prec] = (eng.gcn_features(), eng.pred.clone(), loss.clone(), grads.clone(), hid)
    f0, p0, l0, g0, h0 = out["fp32"]
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))  # 0/0 -> 0 (T = 1: dW_hh = 0)
    lay = unflatten_trainable(g0[0], dims)
    for mode in ("tf32x3", "stepwise"):  # persistent cluster kernels, then the per-step launches
        f1, p1, l1, g1, h1 = out[mode]
        assert rel(f1, f0) <= 2e-5, (mode, "GCN features")
        assert rel(h1, h0) <= 5e-5, (mode, "LSTM hidden states", rel(h1, h0))
        assert rel(p1, p0) <= 1e-4 and rel(l1, l0) <= 1e-4, mode
        for g in range(G):
            a, b = unflatten_trainable(g1[g], dims), unflatten_trainable(g0[g], dims)
            for k in lay:
                assert rel(a[k], b[k]) <= 1e-3, (mode, g, k, rel(a[k], b[k]))
    f1, p1, l1, g1, h1 = out["tf32x3"]
    # and against the CPU oracle for task 0, window 0
    x, y = P.window_xy(feats[0], starts[0][0], T, H)
    l_ref, g_ref, p_ref = P.loss_and_grads(sds[0], x, y, eis[0], T, H, 1.0, 4)
    got = p1[:n].cpu().view(n, H, 12).reshape(-1, 12)
    assert float((got - p_ref).abs().max() / p_ref.abs().max()) <= 1e-4
