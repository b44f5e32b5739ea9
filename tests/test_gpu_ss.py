"""The SS-mode tcgen05 kernels (csrc/wf_gemm_ss.cu) in isolation, every operand layout against a float64 product:
K-major SWIZZLE_64B boxes from row-major planes, K-major and MN-major no-swizzle views of the TB8 layout, MN-major
SWIZZLE_128B boxes from row-major planes; fp16 and bf16 hi/lo planes; resident weight slices of 64 / 128 / 256 columns."""
import pytest
import torch

from weatherforecast_stgcn_maml_b200 import _lib

pytestmark = pytest.mark.gpu


def tile_rows(n):
    return int(_lib.query("wf_tile_rows", n))


def split(x, fmt):
    """fp32 -> (2, n) int16 planes (hi, lo)."""
    x = x.contiguous()
    out = torch.empty(2, x.numel(), dtype=torch.int16, device=x.device)
    _lib.call("wf_split16", _lib.ptr(x), _lib.ptr(out[0]), _lib.ptr(out[1]), x.numel(), fmt, _lib.stream_ptr())
    return out


def joined(planes, fmt):
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    return planes[0].view(dt).double() + planes[1].view(dt).double()


def to_tb8(x, ZT, N, C, fmt):
    """x [ZT, N, C] fp32 -> TB8 planes [2, ZT*tpw*(C/8)*128*8] (padding rows zero) and the value the planes hold."""
    tpw, rpt = (N + 127) // 128, tile_rows(N)
    blocks = torch.zeros(ZT, tpw, C // 8, 128, 8, device=x.device)
    for nt in range(tpw):
        n0, n1 = nt * rpt, min(N, nt * rpt + rpt)
        if n1 > n0:
            blocks[:, nt, :, :n1 - n0] = x[:, n0:n1].reshape(ZT, n1 - n0, C // 8, 8).permute(0, 2, 1, 3)
    planes = split(blocks.reshape(-1), fmt)
    return planes


def from_tb4(c, ZT, N, C):
    """TB4 fp32 [ZT*tpw][C/4][128][4] -> [ZT, N, C]."""
    tpw, rpt = (N + 127) // 128, tile_rows(N)
    t = c.view(ZT, tpw, C // 4, 128, 4).permute(0, 1, 3, 2, 4).reshape(ZT, tpw, 128, C)[:, :, :rpt]
    return t.reshape(ZT, tpw * rpt, C)[:, :N]


@pytest.mark.parametrize("avar", [0, 1])
@pytest.mark.parametrize("bn,K,Ntot,fmt,kp", [(128, 256, 512, 0, 1), (256, 128, 512, 0, 1), (64, 512, 128, 1, 1), (128, 24, 256, 0, 1),
                                              (256, 64, 256, 1, 1), (128, 512, 128, 1, 2), (128, 256, 256, 0, 2)])
@pytest.mark.parametrize("N,T,G,Bw", [(441, 3, 2, 2), (128, 2, 1, 1), (57, 1, 3, 1)])
def test_ss_nodes_gemm_every_a_layout(avar, bn, K, Ntot, fmt, kp, N, T, G, Bw):
    """kp = 2: the K range split over two CTAs per row-tile range, partial products added into the cleared output."""
    if avar == 1 and K % 64:
        pytest.skip("TB8 operands come in whole 64-wide k-blocks (LSTM widths)")
    torch.manual_seed(K + N)
    ZT = G * Bw * T
    scale = 1.0 if fmt == 0 else 1e-3
    x = torch.randn(ZT, N, K, device="cuda") * scale
    W = torch.randn(G, Ntot, K, device="cuda") / K ** 0.5
    b1, b2 = torch.randn(G, Ntot, device="cuda") * scale, torch.randn(G, Ntot, device="cuda") * scale
    w16 = split(W.reshape(-1), fmt)
    if avar == 0:
        a16 = split(x.reshape(-1), fmt)
    else:
        a16 = to_tb8(x, ZT, N, K, fmt)
    tpw = (N + 127) // 128
    C = torch.full((ZT * tpw * Ntot * 128,), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("wf_ss_nodes_gemm", bn, avar, _lib.ptr(a16), a16.shape[1], K, fmt, _lib.ptr(w16[0]), _lib.ptr(w16[1]), Ntot * K,
              Ntot, fmt, _lib.ptr(b1), _lib.ptr(b2), Ntot, _lib.ptr(C), T, N, Bw, G, kp, _lib.ptr(err), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"pipeline error code {int(err.item())}"
    got = from_tb4(C, ZT, N, Ntot).double()
    ref = torch.einsum("gznk,gok->gzno", x.double().view(G, Bw * T, N, K), W.double()) + (b1 + b2).double()[:, None, None, :]
    ref = ref.view(ZT, N, Ntot)
    tol = 5e-6 if fmt == 0 else 5e-5
    assert float((got - ref).abs().max() / ref.abs().max()) <= tol


@pytest.mark.parametrize("N,T,G,Bw", [(441, 4, 2, 1), (128, 3, 1, 2), (57, 2, 3, 1), (300, 1, 1, 1)])
def test_ss_wgrad_tb8_operands_with_step_shift(N, T, G, Bw):
    """[dW_ih | dW_hh] = dG^T [x(t) | h(t-1)] and the bias gradients, both operands MN-major from TB8 planes."""
    torch.manual_seed(N + T)
    ZT, L = G * Bw * T, 128
    dg = torch.randn(ZT, N, 512, device="cuda") * 1e-3
    x = torch.randn(ZT, N, L, device="cuda")
    h = torch.randn(ZT, N, L, device="cuda") * 0.5
    dg16, x16, h16 = to_tb8(dg, ZT, N, 512, 1), to_tb8(x, ZT, N, L, 1), to_tb8(h, ZT, N, L, 1)
    nh = 2 if T > 1 else 1
    part = torch.empty(40 * 513 * 512 * G, device="cuda")
    # the three results live in one per-group buffer (like the flat gradient buffer): one group stride for all of them
    buf = torch.full((G, 2 * 512 * L + 512), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("wf_ss_wgrad", _lib.ptr(dg16), dg16.shape[1], nh, _lib.ptr(x16), x16.shape[1], 0, 0, 0, L, _lib.ptr(h16),
              h16.shape[1], 0, 1, 0, L, T, N, Bw, G, _lib.ptr(part), part.numel(), _lib.ptr(buf), L, L,
              _lib.ptr(buf[0, 512 * L:]) if nh == 2 else None, L, L, _lib.ptr(buf[0, 2 * 512 * L:]), buf.shape[1], _lib.ptr(err),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"pipeline error code {int(err.item())}"
    d0, d1, db = buf[:, :512 * L].view(G, 512, L), buf[:, 512 * L:2 * 512 * L].view(G, 512, L), buf[:, 2 * 512 * L:]
    dgv = dg.double().view(G, Bw, T, N, 512)
    xv, hv = x.double().view(G, Bw, T, N, L), h.double().view(G, Bw, T, N, L)
    r0 = torch.einsum("gwtnm,gwtnc->gmc", dgv, xv)
    rb = dgv.sum(dim=(1, 2, 3))
    assert float((d0.double() - r0).abs().max() / r0.abs().max()) <= 5e-5
    # the row sums ride on a product with exact ones: only dG's own bf16 hi/lo split (~2^-17 per value) is in the way
    assert float((db.double() - rb).abs().max() / rb.abs().max()) <= 5e-5
    if nh == 2:
        r1 = torch.einsum("gwtnm,gwtnc->gmc", dgv[:, :, 1:], hv[:, :, :-1])
        assert float((d1.double() - r1).abs().max() / r1.abs().max()) <= 5e-5


@pytest.mark.parametrize("N,T,G,Bw,F", [(441, 3, 2, 1, 256), (57, 2, 1, 2, 256), (200, 2, 1, 1, 128)])
def test_ss_wgrad_row_major_features(N, T, G, Bw, F):
    """Layer 0: dW_ih = dG^T x with x as row-major bf16 planes (MN-major SWIZZLE_128B boxes), 256 columns per pass."""
    torch.manual_seed(N)
    ZT = G * Bw * T
    dg = torch.randn(ZT, N, 512, device="cuda") * 1e-3
    x = torch.randn(ZT, N, F, device="cuda")
    dg16, x16 = to_tb8(dg, ZT, N, 512, 1), split(x.reshape(-1), 1)
    part = torch.empty(40 * 513 * 512 * G, device="cuda")
    d0 = torch.full((G, 512, F), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for f0 in range(0, F, 256):
        w = min(256, F - f0)
        _lib.call("wf_ss_wgrad", _lib.ptr(dg16), dg16.shape[1], 2 if w > 128 else 1, _lib.ptr(x16), x16.shape[1], 1, 0, f0, F,
                  _lib.ptr(x16), x16.shape[1], 1, 0, f0 + 128, F, T, N, Bw, G, _lib.ptr(part), part.numel(),
                  _lib.ptr(d0[:, :, f0:]), F, w, None, 0, 0, None, 512 * F, _lib.ptr(err), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"pipeline error code {int(err.item())}"
    r0 = torch.einsum("gznm,gznc->gmc", dg.double().view(G, Bw * T, N, 512), x.double().view(G, Bw * T, N, F))
    assert float((d0.double() - r0).abs().max() / r0.abs().max()) <= 5e-5


@pytest.mark.parametrize("cin,R,G,Bw,edges", [(256, 1000, 2, 2, "t0"), (24, 10584, 1, 1, "t0"), (256, 300, 1, 1, "all"),
                                             (64, 129, 1, 3, "none")])
def test_gcn_layer_fwd_ss_vs_fp64(cin, R, G, Bw, edges):
    """One GCN layer on pre-split planes: aggregation of the leading rows + resident-weight GEMM + TMA-store epilogue
    (bias, ReLU), partial last tiles clipped at the window's rows, fp32 input windows (first layer) or fp16 planes."""
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs

    torch.manual_seed(cin + R)
    cout, Z = 256, G * Bw
    nn_ = 60 if edges == "t0" else R
    graphs = []
    for g in range(G):
        if edges == "none":
            ei = torch.zeros(2, 0, dtype=torch.long)
        else:
            ei = torch.randint(0, nn_, (2, 4 * nn_), generator=torch.Generator().manual_seed(g))
        graphs.append(RegionGraph(ei, R, "cuda"))
    sg = StackedGraphs(graphs)
    x = torch.randn(Z, R, cin, device="cuda")
    W = torch.randn(cout, cin, device="cuda") / cin ** 0.5
    b = torch.randn(cout, device="cuda") * 0.1
    w16 = split(W.reshape(-1), 0)
    y16 = torch.full((2, Z * R * cout), 0x7e00, dtype=torch.int16, device="cuda")   # fp16 NaN: unwritten values show
    yb16 = torch.full((2, Z * R * cout), 0x7fc0, dtype=torch.int16, device="cuda")  # bf16 NaN
    xs16 = torch.empty(2, Z * R * cin, dtype=torch.int16, device="cuda")
    agg = int(sg.agg_rows)
    side = torch.empty(2 * Z * max(agg, 1) * cin, dtype=torch.int16, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for as_planes in (False, True):
        x16 = split(x.reshape(-1), 0) if as_planes else None
        _lib.call("wf_gcn_layer_fwd_ss", None if as_planes else _lib.ptr(x), None, _lib.ptr(x16), _lib.ptr(xs16),
                  _lib.ptr(w16[0]), _lib.ptr(w16[1]), _lib.ptr(b), _lib.ptr(sg.rowptr), _lib.ptr(sg.col), _lib.ptr(sg.val),
                  sg.rowptr_stride, sg.csr_stride, agg, _lib.ptr(side) if agg else None, R, cin, cout, G, Bw, 1, _lib.ptr(y16),
                  _lib.ptr(yb16) if as_planes else None, 0.0, None, 0, _lib.ptr(err), _lib.stream_ptr())
        torch.cuda.synchronize()
        assert int(err.item()) == 0, f"error code {int(err.item())}"
        got = joined(y16, 0).view(Z, R, cout)
        ref = torch.empty(Z, R, cout, dtype=torch.float64, device="cuda")
        for z in range(Z):
            gr = graphs[z // Bw]
            rp = gr.rowptr.long()
            rows = torch.repeat_interleave(torch.arange(R, device="cuda"), rp[1:] - rp[:-1])
            A = torch.zeros(R, R, dtype=torch.float64, device="cuda")
            A.index_put_((rows, gr.col[:rows.numel()].long()), gr.val[:rows.numel()].double(), accumulate=True)
            ref[z] = torch.relu((A @ x[z].double()) @ W.double().t() + b.double())
        assert float((got - ref).abs().max() / ref.abs().max()) <= 1e-5
        if as_planes:  # the same values once more as bf16 hi/lo planes (2^-17 per value)
            assert float((joined(yb16, 1).view(Z, R, cout) - got).abs().max() / ref.abs().max()) <= 2e-5
        y16.fill_(0x7e00)
