"""loss.backward() through GCNConv + ReLU (+ Dropout) on the tensor cores (wf_gcn_layer_bwd_ss, through the C ABI)
against a float64 restatement of model.py:31-42 under autograd: dW = dZ^T (A_hat X), db = dZ^T 1, dX = A_hat^T (dZ W)
with dZ = dY * mask * (Y > 0).  Gradient tolerance (north_star): 1e-3 relative; measured ~1e-5 (bf16 hi/lo splits)."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _dense_ahat(graph, R):
    rp, col, val = graph.rowptr.cpu(), graph.col.cpu(), graph.val.cpu()
    A = torch.zeros(R, R, dtype=torch.float64)
    for r in range(R):
        for p in range(int(rp[r]), int(rp[r + 1])):
            A[r, int(col[p])] += float(val[p])
    return A


@pytest.mark.parametrize("p_drop", [0.0, 0.3])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("cin,cout,need_x", [(256, 256, True), (128, 256, True), (24, 256, False), (256, 128, True), (64, 384, False)])
@pytest.mark.parametrize("nlat,nlon,T,bw", [(5, 7, 6, 2), (12, 13, 3, 1), (3, 3, 2, 3)])
def test_gcn_layer_bwd_ss(nlat, nlon, T, bw, cin, cout, need_x, relu, p_drop):
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph

    torch.manual_seed(cin + cout + T)
    lats, lons = synth.region_grid(nlat, nlon)
    ei = P.knn_edges_ckdtree(lats, lons, 4)
    n = nlat * nlon
    R = T * n                                   # time-major rows; only the t = 0 slice has neighbours
    g = RegionGraph(ei, R, "cuda")
    dev = "cuda"
    x = torch.randn(bw * R, cin, device=dev)
    W = torch.randn(cout, cin, device=dev) / cin ** 0.5
    y = torch.randn(bw * R, cout, device=dev)   # stands for the forward output: only its sign pattern is read
    dy = torch.randn(bw * R, cout, device=dev) * 1e-3
    rng = torch.tensor([7, 3], dtype=torch.int64, device=dev)
    site = 2
    dx = torch.full((bw * R, cin), float("nan"), device=dev) if need_x else None
    dw = torch.full((cout, cin), float("nan"), device=dev)
    db = torch.full((cout,), float("nan"), device=dev)
    nbytes = int(_lib.query("wf_gcn_layer_bwd_ss_workspace_bytes", R, cin, cout, bw))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("wf_gcn_layer_bwd_ss", _lib.ptr(x), _lib.ptr(y), _lib.ptr(dy), _lib.ptr(W), _lib.ptr(g.rowptr), _lib.ptr(g.col),
              _lib.ptr(g.val), _lib.ptr(g.rowptr_t), _lib.ptr(g.col_t), _lib.ptr(g.val_t), R, cin, cout, bw, int(relu),
              float(p_drop), _lib.ptr(rng), site, _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), nbytes, _lib.ptr(err),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    # float64 reference
    A = _dense_ahat(g, R)
    mask = torch.ones(bw * R, cout, dtype=torch.float64)
    if p_drop > 0:
        ones = torch.ones(bw * R, cout, device=dev)
        m = torch.empty_like(ones)
        _lib.call("wf_dropout_apply", _lib.ptr(ones), 0, bw * R, cout, bw * R, cout, float(p_drop), _lib.ptr(rng), site, _lib.ptr(m),
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        mask = m.cpu().double()
    dz = dy.cpu().double() * mask
    if relu:
        dz = dz * (y.cpu() > 0).double()
    xd, Wd = x.cpu().double(), W.cpu().double()
    ax = torch.cat([A @ xd[z * R:(z + 1) * R] for z in range(bw)])
    assert rel_err(dw, dz.t() @ ax) <= 1e-4
    assert rel_err(db, dz.sum(0)) <= 1e-4
    if need_x:
        pz = dz @ Wd
        ref_dx = torch.cat([A.t() @ pz[z * R:(z + 1) * R] for z in range(bw)])
        assert rel_err(dx, ref_dx) <= 1e-4


def test_gcn_layer_bwd_ss_refuses_unsupported_widths():
    dev = "cuda"
    t = torch.zeros(16, device=dev)
    for cin, cout, with_dx in ((256, 96, False), (20, 128, False), (512, 128, False), (64, 128, True)):
        with pytest.raises(RuntimeError):
            _lib.call("wf_gcn_layer_bwd_ss", _lib.ptr(t), _lib.ptr(t), _lib.ptr(t), _lib.ptr(t), None, None, None, None, None, None, 8, cin,
                      cout, 1, 0, 0.0, None, 0, _lib.ptr(t) if with_dx else None, _lib.ptr(t), _lib.ptr(t), _lib.ptr(t), 1 << 20, None,
                      _lib.stream_ptr())
