"""GPU parity of the model path through the C ABI: GCN layers, STGCN fwd/bwd, hybrid forward,
MSE, BPTT -- against the oracle and the fixtures frozen from the reference.
Tolerances (north_star): 1e-4 relative forward, 1e-3 relative gradients."""
import numpy as np
import pytest
import torch

from conftest import check_summary, golden_case, rel_err
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 1e-3


@pytest.fixture(params=["tf32x3", "fp32"])
def precision(request):
    """The drop-in modules on both kernel families: tcgen05 (16-bit hi/lo operand splits, what the benchmark runs,
    taken whenever the shapes allow) and exact FP32 CUDA cores."""
    from weatherforecast_stgcn_maml_b200 import functional as WF

    WF.set_precision(request.param)
    yield request.param
    WF.check()
    WF.set_precision("tf32x3")


def _models(cfg, sd, dropout=0.0):
    from weatherforecast_stgcn_maml_b200.hybrid_model import HybridSTGCN_LSTM
    from weatherforecast_stgcn_maml_b200.model import STGCN

    base = STGCN(cfg["cin"], cfg["hidden"], out_channels=cfg["out"], window_size=cfg["T"],
                 forecast_horizon=cfg["H"], dropout_rate=dropout)
    hyb = HybridSTGCN_LSTM(base, lstm_hidden_size=cfg["L"], lstm_num_layers=cfg["layers"], lstm_dropout=dropout,
                           out_channels=cfg["out"], forecast_horizon=cfg["H"], freeze_base=False)
    res = hyb.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return hyb.cuda()


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("cin,cout", [(24, 256), (256, 256), (24, 32), (40, 72)])
def test_gcn_layer_forward(relu, cin, cout, precision):
    from weatherforecast_stgcn_maml_b200 import functional as WF
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph

    torch.manual_seed(0)
    lats, lons = synth.region_grid(5, 7)
    ei = P.knn_edges_ckdtree(lats, lons, 4)
    T, n = 6, 35
    x = torch.randn(T * n, cin)
    W = torch.randn(cout, cin) / cin ** 0.5
    b = torch.randn(cout) * 0.1
    ref = P.gcn_conv(x, ei, W, b)
    ref = torch.relu(ref) if relu else ref
    g = RegionGraph(ei, T * n, "cuda")
    got = WF.gcn_conv(x.cuda(), W.cuda(), b.cuda(), g, relu=relu)
    tol = 1e-5 if precision == "fp32" or cout % 128 else 2e-5  # fp16 hi/lo operand split: ~2^-20 per product
    assert rel_err(got, ref) <= tol
    # rows >= N see only their self loop: out = x W^T + b
    lin = x[n:] @ W.t() + b
    lin = torch.relu(lin) if relu else lin
    assert rel_err(got[n:], lin) <= tol


def test_gcn_layer_edges_at_every_row():
    """A general graph over all R rows (not the reference's t=0-only quirk) also aggregates correctly."""
    from weatherforecast_stgcn_maml_b200 import functional as WF
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph

    torch.manual_seed(1)
    R, cin, cout = 300, 24, 64
    ei = torch.randint(0, R, (2, 1500))
    x, W, b = torch.randn(R, cin), torch.randn(cout, cin) * 0.2, torch.randn(cout) * 0.1
    got = WF.gcn_conv(x.cuda(), W.cuda(), b.cuda(), RegionGraph(ei, R, "cuda"), relu=False)
    assert rel_err(got, P.gcn_conv(x, ei, W, b)) <= 1e-5


@pytest.mark.parametrize("name", ["hybrid_small", "hybrid_v5_k4", "hybrid_v5_k8"])
def test_hybrid_forward_and_backward_vs_reference_fixture(name, precision):
    z, cfg, sd, feats, ei = golden_case(name)
    T, H = cfg["T"], cfg["H"]
    hyb = _models(cfg, sd)
    hyb.train()
    x, y = P.window_xy(feats, 0, T, H)
    bf = hyb.extract_base_features(x.cuda(), ei.cuda())
    assert not bf.requires_grad
    check_summary(bf, z["base_features_summary"], z["base_features_samples"], FWD_TOL, "base features")
    pred = hyb(x.cuda(), ei.cuda())
    assert tuple(pred.shape) == (H * cfg["nlat"] * cfg["nlon"], cfg["out"])
    assert rel_err(pred, torch.from_numpy(z["pred"])) <= FWD_TOL
    loss = torch.nn.MSELoss()(pred, y.cuda())
    assert abs(loss.item() - float(z["loss"])) <= FWD_TOL * float(z["loss"])
    loss.backward()
    for k_, p in hyb.named_parameters():
        if k_.startswith("base_stgcn."):
            assert p.grad is None, k_
        else:
            check_summary(p.grad, z[f"grad_summary/{k_}"], z[f"grad_samples/{k_}"], GRAD_TOL, k_)
            if f"grad/{k_}" in z.files:
                assert rel_err(p.grad, torch.from_numpy(z[f"grad/{k_}"])) <= GRAD_TOL, k_


@pytest.mark.parametrize("name", ["hybrid_small", "hybrid_v5_k4"])
def test_stgcn_forward_backward_vs_reference_fixture(name, precision):
    """model.py:30-52 end to end: the only differentiable use of the graph convolution (D4)."""
    from weatherforecast_stgcn_maml_b200.model import STGCN

    z, cfg, sd, feats, ei = golden_case(name)
    T, H = cfg["T"], cfg["H"]
    base = STGCN(cfg["cin"], cfg["hidden"], out_channels=cfg["out"], window_size=T, forecast_horizon=H,
                 dropout_rate=0.0)
    base.load_state_dict({k[len("base_stgcn."):]: v for k, v in sd.items() if k.startswith("base_stgcn.")})
    base = base.cuda().train()
    x, y = P.window_xy(feats, 0, T, H)
    xs = x.cuda().requires_grad_(True)
    pred = base(xs, ei.cuda())
    assert rel_err(pred, torch.from_numpy(z["stgcn_pred"])) <= FWD_TOL
    loss = torch.nn.MSELoss()(pred, y.cuda())
    assert abs(loss.item() - float(z["stgcn_loss"])) <= FWD_TOL * float(z["stgcn_loss"])
    loss.backward()
    for k_, p in base.named_parameters():
        check_summary(p.grad, z[f"stgcn_grad_summary/{k_}"], z[f"stgcn_grad_samples/{k_}"], GRAD_TOL, k_)
    check_summary(xs.grad, z["stgcn_dx_summary"], z["stgcn_dx_samples"], GRAD_TOL, "dx")


def test_stgcn_backward_general_graph_vs_oracle(precision):
    """dX = A_hat^T path with real neighbour mixing in every layer (window_size = 1 -> all rows have edges)."""
    from weatherforecast_stgcn_maml_b200.model import STGCN

    torch.manual_seed(3)
    lats, lons = synth.region_grid(6, 5)
    ei = P.knn_edges_ckdtree(lats, lons, 4)
    n = 30
    sd = synth.init_v5_state_dict(11, gcn_bias_scale=0.05, hidden=64, lstm_hidden=32, lstm_layers=1, horizon=2)
    base_sd = {k[len("base_stgcn."):]: v for k, v in sd.items() if k.startswith("base_stgcn.")}
    x, y = torch.randn(n, 24), torch.randn(2 * n, 12)
    leaf = {k: v.clone().requires_grad_(True) for k, v in base_sd.items()}
    xr = x.clone().requires_grad_(True)
    pr = P.stgcn_forward(leaf, xr, ei, 1, 2, 12)
    gr = torch.autograd.grad(torch.nn.functional.mse_loss(pr, y), list(leaf.values()) + [xr])
    base = STGCN(24, 64, out_channels=12, window_size=1, forecast_horizon=2, dropout_rate=0.0)
    base.load_state_dict(base_sd)
    base = base.cuda()
    xs = x.cuda().requires_grad_(True)
    pred = base(xs, ei.cuda())
    assert rel_err(pred, pr) <= FWD_TOL
    torch.nn.functional.mse_loss(pred, y.cuda()).backward()
    for (k_, _), g in zip(leaf.items(), gr):
        assert rel_err(dict(base.named_parameters())[k_].grad, g) <= GRAD_TOL, k_
    assert rel_err(xs.grad, gr[-1]) <= GRAD_TOL


def test_engine_batched_windows_equal_per_window_oracle():
    """Batch B = B independent windows with batch-1 semantics, two tasks with
    different graphs and different fast weights in one launch."""
    from weatherforecast_stgcn_maml_b200.engine import (HybridEngine, V5Dims, flatten_trainable,
                                                        gcn_weights_from_state_dict, unflatten_trainable)
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs

    dims = V5Dims(num_nodes=20, window=5, horizon=3, hidden=32, lstm_hidden=32, lstm_layers=2)
    G, Bw = 2, 3
    lats, lons = synth.region_grid(4, 5)
    eis = [P.knn_edges_ckdtree(lats, lons, 4), P.knn_edges_ckdtree(lats, lons, 3)]
    eis[1] = torch.cat([eis[1], eis[1][:, :20]], 1)  # same capacity (80 edges), duplicate edges allowed
    assert eis[0].shape == eis[1].shape
    base = synth.init_v5_state_dict(5, gcn_bias_scale=0.05, hidden=32, lstm_hidden=32, lstm_layers=2, horizon=3)
    sds = [base, {k: (v + 0.01 * torch.randn_like(v) if k.startswith(("lstm.", "output_layer.")) else v)
                  for k, v in base.items()}]
    feats = torch.stack([synth.synth_features(20, 20, 100 + g) for g in range(G)])
    starts = [[0, 3, 5], [1, 2, 6]]
    dev = "cuda"
    eng = HybridEngine(dims, G, Bw, dev)
    per, per_task = 20 * 24, 20 * 20 * 24
    xo = torch.tensor([g * per_task + s * per for g in range(G) for s in starts[g]], device=dev)
    to = xo + (dims.window + 1) * per
    theta = torch.stack([flatten_trainable(sd, dims) for sd in sds]).to(dev)
    graphs = StackedGraphs([RegionGraph(ei, dims.R, dev) for ei in eis])
    fd = feats.to(dev)
    loss, grads = eng.forward_backward(fd, 24, 0, xo, gcn_weights_from_state_dict(base, dev), graphs, theta, eng.P,
                                       feat=fd, tgt_off=to, feat_ld=24, grad_scale=0.5)
    torch.cuda.synchronize()
    for g in range(G):
        gsum = None
        for b, s in enumerate(starts[g]):
            x, y = P.window_xy(feats[g], s, dims.window, dims.horizon)
            l_ref, g_ref, p_ref = P.loss_and_grads(sds[g], x, y, eis[g], dims.window, dims.horizon, 0.5, 2)
            w = g * Bw + b
            got = eng.pred[w * 20:(w + 1) * 20].cpu().view(20, 3, 12).reshape(-1, 12)
            assert rel_err(got, p_ref) <= FWD_TOL
            assert abs(loss[w].item() - float(l_ref) / 0.5) <= FWD_TOL * float(l_ref) / 0.5
            gsum = g_ref if gsum is None else {k: gsum[k] + g_ref[k] for k in g_ref}
        got_g = unflatten_trainable(grads[g].cpu(), dims)
        for k in gsum:  # gradients of windows of one task add up (sum of per-window losses)
            assert rel_err(got_g[k], gsum[k]) <= GRAD_TOL, (g, k)


def test_mse_explicit_targets_match_inplace_targets():
    from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims

    dims = V5Dims(num_nodes=12, window=4, horizon=2, hidden=32, lstm_hidden=32, lstm_layers=1)
    eng = HybridEngine(dims, 1, 2, "cuda")
    feats = synth.synth_features(12, 12, 9).cuda()
    eng.pred.normal_()
    per = 12 * 24
    to = torch.tensor([(0 + 5) * per, (2 + 5) * per], device="cuda")
    eng.mse(feat=feats, tgt_off=to, feat_ld=24)
    l1, d1 = eng.loss.clone(), eng.dpred.clone()
    y = torch.stack([P.window_xy(feats.cpu(), s, 4, 2)[1].reshape(-1) for s in (0, 2)]).cuda()
    eng.mse(y=y)
    assert torch.equal(l1, eng.loss) and torch.equal(d1, eng.dpred)
    ref = ((eng.pred.view(2, -1) - y) ** 2).mean(1)
    assert rel_err(eng.loss, ref) <= 1e-5


def test_state_dict_roundtrip_and_deepcopy():
    import copy

    z, cfg, sd, feats, ei = golden_case("hybrid_small")
    hyb = _models(cfg, sd)
    out = hyb.state_dict()
    assert list(out.keys()) == list(sd.keys())
    assert all(torch.equal(out[k].cpu(), sd[k]) for k in sd)
    hyb.lstm.flatten_parameters()
    clone = copy.deepcopy(hyb)
    x, _ = P.window_xy(feats, 0, cfg["T"], cfg["H"])
    assert torch.equal(clone(x.cuda(), ei.cuda()), hyb(x.cuda(), ei.cuda()))
    assert len(hyb.get_trainable_parameters()) == 4 * cfg["layers"] + 2
    hyb.freeze_base_model()
    assert not any(p.requires_grad for p in hyb.base_stgcn.parameters())
    hyb.unfreeze_base_model()
    assert all(p.requires_grad for p in hyb.base_stgcn.parameters())


def test_forward_rejects_cpu_tensors():
    z, cfg, sd, feats, ei = golden_case("hybrid_small")
    hyb = _models(cfg, sd)
    x, _ = P.window_xy(feats, 0, cfg["T"], cfg["H"])
    with pytest.raises(RuntimeError):
        hyb(x, ei)


def test_side_stream_overlap_is_bit_identical_to_one_stream():
    """The engine forks the weight staging and the head's parameter gradients onto a side stream (also inside captured
    graphs); the result must not depend on it: same bits with overlap on and off, eager and replayed from a CUDA graph."""
    from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable, gcn_weights_from_state_dict
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs

    dims = V5Dims(num_nodes=35, window=6, horizon=3)
    G, Bw, dev = 2, 2, "cuda"
    lats, lons = synth.region_grid(5, 7)
    ei = P.knn_edges_ckdtree(lats, lons, 4)
    sd = synth.init_v5_state_dict(11, gcn_bias_scale=0.05, horizon=3)
    feats = torch.stack([synth.synth_features(16, 35, 300 + g) for g in range(G)]).to(dev)
    per, per_task = 35 * 24, 16 * 35 * 24
    xo = torch.tensor([g * per_task + b * per for g in range(G) for b in range(Bw)], device=dev)
    to = xo + (dims.window + 1) * per
    theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).to(dev)
    graphs = StackedGraphs([RegionGraph(ei, dims.R, dev) for _ in range(G)])
    gw = gcn_weights_from_state_dict(sd, dev)
    out = {}
    for overlap in (True, False):
        eng = HybridEngine(dims, G, Bw, dev)
        if not eng.seq:
            pytest.skip("persistent tensor-core path not available for this shape")
        eng.overlap = overlap
        run = lambda: eng.forward_backward(feats, 24, 0, xo, gw, graphs, theta, eng.P, feat=feats, tgt_off=to, feat_ld=24)
        run()
        torch.cuda.synchronize()
        eng.check()
        out[overlap, "eager"] = (eng.loss.clone(), eng.grads.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        eng.grads.zero_()
        graph.replay()
        torch.cuda.synchronize()
        eng.check()
        out[overlap, "graph"] = (eng.loss.clone(), eng.grads.clone())
    ref = out[False, "eager"]
    for key, val in out.items():
        assert torch.equal(val[0], ref[0]) and torch.equal(val[1], ref[1]), key
    assert float(ref[1].abs().max()) > 0
