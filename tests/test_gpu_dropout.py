"""Train-mode dropout at the reference's three sites (hybrid_model.py:47,58,67-73,108; model.py:27,33-42).

torch's generator stream cannot be matched by kernels that batch nodes, windows and tasks, so the
contract tested here is: (i) the masks have the right distribution (keep rate 1 - p, survivors scaled by 1/(1-p),
independent across sites / passes / seeds); (ii) forward AND backward use the same mask at exactly the reference's
sites -- checked by reading the masks back (wf_dropout_apply on ones) and comparing predictions, loss and every
gradient with the CPU oracle run on those masks, at the north_star tolerances (1e-4 forward, 1e-3 gradients);
(iii) p = 0 / eval mode is bit-identical to the deterministic path; (iv) the reference's training loop shape
(model built with dropout_rate=0.2, lstm_dropout=0.2, .train()) runs unedited on the drop-in modules.
"""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import _lib, synth

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 1e-3
SITE_LSTM, SITE_HEAD = 16, 32


def read_mask(seed, counter, site, rows, cols, p, dev="cuda"):
    """The scaled keep-mask of a site for (seed, pass counter), via the library's own stand-alone pass."""
    ones = torch.ones(rows, cols, device=dev)
    out = torch.empty_like(ones)
    rng = torch.tensor([int(seed), int(counter)], dtype=torch.int64, device=dev)
    _lib.call("wf_dropout_apply", _lib.ptr(ones), 0, rows, cols, rows, cols, float(p), _lib.ptr(rng), int(site),
              _lib.ptr(out), _lib.stream_ptr())
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("p", [0.2, 0.5, 0.05])
def test_mask_distribution_scaling_and_independence(p):
    rows, cols = 4096, 256
    m = read_mask(42, 0, 1, rows, cols, p)
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0] == 0.0 and abs(vals[1].item() - 1.0 / (1.0 - p)) < 1e-6  # survivors scaled 1/(1-p)
    n = rows * cols
    keep = (m > 0).double().mean().item()
    assert abs(keep - (1 - p)) < 5 * np.sqrt(p * (1 - p) / n)  # keep rate 1 - p (5 sigma)
    assert abs(m.mean().item() - 1.0) < 5 * np.sqrt(p / (1 - p) / n)  # E[mask] = 1: activations keep their expectation
    # per-row and per-column rates are flat as well (no structure along either axis)
    assert ((m > 0).double().mean(0) - (1 - p)).abs().max() < 6 * np.sqrt(p * (1 - p) / rows)
    assert ((m > 0).double().mean(1) - (1 - p)).abs().max() < 6 * np.sqrt(p * (1 - p) / cols)
    assert torch.equal(m, read_mask(42, 0, 1, rows, cols, p))  # a pure function of (seed, pass, site, element)
    for other in (read_mask(43, 0, 1, rows, cols, p), read_mask(42, 1, 1, rows, cols, p), read_mask(42, 0, 2, rows, cols, p),
                  read_mask(42, 1 << 32, 1, rows, cols, p)):
        agree = ((m > 0) == (other > 0)).double().mean().item()
        assert abs(agree - (p * p + (1 - p) ** 2)) < 5 / np.sqrt(n) + 1e-3  # independent draws
    # the same element keeps its decision when the tensor is read through a strided gather (head site addressing)
    src = torch.ones(3 * rows, cols, device="cuda")
    out = torch.empty(rows, cols, device="cuda")
    rng = torch.tensor([42, 0], dtype=torch.int64, device="cuda")
    _lib.call("wf_dropout_apply", _lib.ptr(src[2:]), 3 * 8 * cols, 8, cols, rows, cols, float(p), _lib.ptr(rng), 1,
              _lib.ptr(out), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), m)


def test_p_zero_and_null_rng_are_copies():
    x = torch.randn(64, 32, device="cuda")
    out = torch.empty_like(x)
    rng = torch.tensor([1, 2], dtype=torch.int64, device="cuda")
    _lib.call("wf_dropout_apply", _lib.ptr(x), 0, 64, 32, 64, 32, 0.0, _lib.ptr(rng), 0, _lib.ptr(out), _lib.stream_ptr())
    assert torch.equal(out, x)
    _lib.call("wf_dropout_apply", _lib.ptr(x), 0, 64, 32, 64, 32, 0.3, None, 0, _lib.ptr(out), _lib.stream_ptr())
    assert torch.equal(out, x)
    with pytest.raises(RuntimeError):
        _lib.call("wf_dropout_apply", _lib.ptr(x), 0, 64, 32, 64, 32, 1.0, _lib.ptr(rng), 0, _lib.ptr(out), _lib.stream_ptr())


def _case(nlat=6, nlon=7, T=6, H=2, G=2, seed=3, k=4):
    from weatherforecast_stgcn_maml_b200.engine import V5Dims

    dims = V5Dims(num_nodes=nlat * nlon, window=T, horizon=H)
    lats, lons = synth.region_grid(nlat, nlon)
    ei = P.knn_edges_canonical(lats, lons, k)
    sd = synth.init_v5_state_dict(seed, gcn_bias_scale=0.05, in_channels=24, hidden=dims.hidden,
                                  lstm_hidden=dims.lstm_hidden, lstm_layers=dims.lstm_layers, horizon=H)
    feats = [synth.synth_features(T + H + 4, dims.num_nodes, 5 + g, synth.koppen_table(3)[2]) for g in range(G)]
    return dims, ei, sd, feats


def _site_masks(dims, Z, seed, counter, ps):
    """Masks of every site for pass ``counter`` in the oracle's layouts, per window z."""
    d = dims
    R, N, T, L, Ls = d.R, d.num_nodes, d.window, d.lstm_hidden, d.lstm_layers
    gcn = [read_mask(seed, counter, i, Z * R, d.hidden, ps[0]) if ps[0] > 0 else None for i in range(3)]
    lstm = [read_mask(seed, counter, SITE_LSTM + l, Z * T * N, L, ps[1]).view(Z, T, N, L) if ps[1] > 0 else None
            for l in range(Ls - 1)]
    head = read_mask(seed, counter, SITE_HEAD, Z * N, L, ps[2]) if ps[2] > 0 else None
    out = []
    for z in range(Z):
        out.append({"gcn": [None if m is None else m[z * R:(z + 1) * R] for m in gcn],
                    "lstm": [None if m is None else m[z].permute(1, 0, 2) for m in lstm],
                    "head": None if head is None else head[z * N:(z + 1) * N]})
    return out


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
@pytest.mark.parametrize("ps", [(0.2, 0.2, 0.2), (0.0, 0.3, 0.0), (0.25, 0.0, 0.0), (0.0, 0.0, 0.4)])
def test_engine_forward_backward_use_the_oracles_masks(precision, ps):
    """Every site on and each site alone: predictions, loss and all 18 gradients equal the oracle's on the SAME masks,
    for two tasks with their own fast weights in one launch, over two consecutive passes (fresh masks each)."""
    from weatherforecast_stgcn_maml_b200.engine import (HybridEngine, flatten_trainable, gcn_weights_from_state_dict,
                                                        unflatten_trainable)
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph

    dims, ei, sd, feats = _case()
    G, dev, seed = 2, "cuda", 1234
    sds = [sd, {k: (v + 0.02 * torch.randn(v.shape, generator=torch.Generator().manual_seed(7))
                    if k.startswith(("lstm.", "output_layer.")) else v) for k, v in sd.items()}]
    eng = HybridEngine(dims, G, 1, dev, precision=precision, dropout=ps, seed=seed)
    assert eng.stochastic and eng.seq == (precision == "tf32x3")
    graph = RegionGraph(ei, dims.R, dev)
    from weatherforecast_stgcn_maml_b200.graph import StackedGraphs
    graphs = StackedGraphs([graph, RegionGraph(ei, dims.R, dev)])
    fd = torch.stack(feats).to(dev)
    per, per_task = dims.num_nodes * dims.in_channels, feats[0].numel()
    theta = torch.stack([flatten_trainable(s_, dims) for s_ in sds]).to(dev)
    gw = gcn_weights_from_state_dict(sd, dev)
    preds = []
    for it, start in enumerate((1, 2)):
        xo = torch.tensor([g * per_task + start * per for g in range(G)], dtype=torch.long, device=dev)
        to = xo + (dims.window + 1) * per
        masks = _site_masks(dims, G, seed, it, ps)  # the pass counter starts at 0 and moves by one per forward+backward
        loss, grads = eng.forward_backward(fd, dims.in_channels, 0, xo, gw, graphs, theta, eng.P, feat=fd, tgt_off=to,
                                           feat_ld=dims.in_channels)
        torch.cuda.synchronize()
        eng.check()
        assert int(eng.rng[1].item()) == it + 1
        preds.append(eng.pred.clone())
        for g in range(G):
            x, y = P.window_xy(feats[g], start, dims.window, dims.horizon)
            l_ref, g_ref, p_ref = P.loss_and_grads(sds[g], x, y, ei, dims.window, dims.horizon, 1.0, dims.lstm_layers,
                                                   masks=masks[g])
            got = eng.pred[g * dims.num_nodes:(g + 1) * dims.num_nodes].cpu().view(dims.num_nodes, dims.horizon, 12).reshape(-1, 12)
            assert rel_err(got, p_ref) <= FWD_TOL, (it, g)
            assert abs(loss[g].item() - float(l_ref)) <= FWD_TOL * float(l_ref)
            got_g = unflatten_trainable(grads[g].cpu(), dims)
            for name, ref in g_ref.items():
                assert rel_err(got_g[name], ref) <= GRAD_TOL, (it, g, name)
    # eval mode: bit-identical to an engine built without dropout
    eng.eval()
    plain = HybridEngine(dims, G, 1, dev, precision=precision)
    xo = torch.tensor([g * per_task + per for g in range(G)], dtype=torch.long, device=dev)
    to = xo + (dims.window + 1) * per
    l1, g1 = eng.forward_backward(fd, dims.in_channels, 0, xo, gw, graphs, theta, eng.P, feat=fd, tgt_off=to, feat_ld=dims.in_channels)
    l1, g1, p1 = l1.clone(), g1.clone(), eng.pred.clone()
    l2, g2 = plain.forward_backward(fd, dims.in_channels, 0, xo, gw, graphs, theta, plain.P, feat=fd, tgt_off=to, feat_ld=dims.in_channels)
    assert torch.equal(p1, plain.pred) and torch.equal(l1, l2) and torch.equal(g1, g2)
    assert int(eng.rng[1].item()) == 2  # no pass consumed in eval mode
    assert not torch.equal(p1, preds[0])


def test_keep_rate_and_scaling_inside_the_fused_kernels():
    """Site by site on the tensor-core path: what the fused epilogue / recurrence kernel leaves in the activations."""
    from weatherforecast_stgcn_maml_b200.engine import HybridEngine, flatten_trainable, gcn_weights_from_state_dict
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph

    dims, ei, sd, feats = _case(G=1)
    dev, p = "cuda", 0.2
    fd = feats[0].to(dev)
    per = dims.num_nodes * dims.in_channels
    xo = torch.tensor([per], dtype=torch.long, device=dev)
    graph = RegionGraph(ei, dims.R, dev)
    gw = gcn_weights_from_state_dict(sd, dev)
    theta = flatten_trainable(sd, dims).to(dev)
    on = HybridEngine(dims, 1, 1, dev, dropout=(p, p, p), seed=9, keep_gcn_activations=True)
    off = HybridEngine(dims, 1, 1, dev, keep_gcn_activations=True)
    on.gcn_forward(fd, dims.in_channels, 0, xo, gw, graph)
    off.gcn_forward(fd, dims.in_channels, 0, xo, gw, graph)
    torch.cuda.synchronize()
    # GCN layer 1: relu(conv1) * mask; the mask read back reproduces it from the deterministic output (the planes hold
    # split(y * m) against split(y) * m: equal to the 2^-22 of the hi/lo split)
    m0 = read_mask(9, 0, 0, dims.R, dims.hidden, p).to(dev)
    a_on, a_off = on.gcn_features(0), off.gcn_features(0)
    assert torch.allclose(a_on, a_off * m0, rtol=1e-6, atol=1e-7)
    assert torch.equal(a_on == 0, (a_off * m0) == 0)
    pos = a_off > 0
    kept = (a_on[pos] != 0).double().mean().item()
    assert abs(kept - (1 - p)) < 5 * np.sqrt(p * (1 - p) / int(pos.sum()))
    assert torch.allclose(a_on[pos & (m0 > 0)], a_off[pos & (m0 > 0)] / (1 - p), rtol=1e-5, atol=1e-7)
    # LSTM layer 0 output as the next layer reads it (TB4, masked) vs the plain recurrence fed the same features
    on.lstm_head_forward(theta, 0, feats=off.feats)
    off.lstm_head_forward(theta, 0, feats=off.feats)
    torch.cuda.synchronize()
    h_on, h_off = on.hidden_states()[0], off.hidden_states()[0]
    ml = read_mask(9, 0, SITE_LSTM, dims.R, dims.lstm_hidden, p).to(dev)
    assert torch.equal(h_on, h_off)  # layer 0 itself recurs on the unmasked h (and dW_hh reads it): bit-identical
    hm = on.hidden_states(masked=True)[0]  # ... the NEXT layer reads the masked copy (fp16 hi/lo planes)
    h16_off = off.hidden_states(masked=True)[0]  # the same planes without a mask (h_off is read from the bf16 pair: 2^-17)
    assert torch.allclose(hm, h16_off * ml, rtol=1e-6, atol=1e-7) and torch.equal(hm == 0, (h16_off * ml) == 0)
    assert torch.allclose(hm, h_off * ml, rtol=2e-5, atol=1e-6)
    on.check(); off.check()


def test_module_api_trains_like_the_reference_inner_loop():
    """Zero-edit route: install_dropin_modules(), then the reference's inner_loop_v4 body verbatim in shape
    (train_hybrid_maml_v5.py:110-141) on a model built as the reference builds it (:191-211: dropout_rate=0.2,
    lstm_dropout=0.2, freeze_base=False) -- deepcopy, .train(), SGD over ALL parameters, MSE, backward, clip, step."""
    import weatherforecast_stgcn_maml_b200 as wf

    saved = {k: sys.modules.get(k) for k in ("model", "hybrid_model", "graphBuilder", "embed_utils", "dataset", "adaptive_scheduler")}
    wf.install_dropin_modules()
    try:
        from embed_utils import KoppenEmbedding
        from hybrid_model import HybridSTGCN_LSTM
        from model import STGCN

        torch.manual_seed(0)
        dims, ei, sd, feats = _case(G=1)
        dev = torch.device("cuda")
        base_stgcn = STGCN(in_channels=24, hidden_channels=256, out_channels=12, window_size=dims.window,
                           forecast_horizon=dims.horizon, dropout_rate=0.2).to(dev)
        hybrid_model = HybridSTGCN_LSTM(base_stgcn=base_stgcn, lstm_hidden_size=128, lstm_num_layers=4, lstm_dropout=0.2,
                                        out_channels=12, forecast_horizon=dims.horizon, freeze_base=False).to(dev)
        hybrid_model.load_state_dict(sd)
        koppen_embed = KoppenEmbedding(embedding_dim=8).to(dev)
        hybrid_model.lstm.flatten_parameters()
        before = {k: v.detach().clone() for k, v in hybrid_model.state_dict().items()}

        temp_model = copy.deepcopy(hybrid_model)
        temp_koppen = copy.deepcopy(koppen_embed)
        temp_model.train()
        temp_koppen.train()
        optimizer = torch.optim.SGD(list(temp_model.parameters()) + list(temp_koppen.parameters()), lr=0.01)
        criterion = torch.nn.MSELoss()
        losses = []
        for epoch in range(2):
            for batch_idx in range(3):
                x, y = P.window_xy(feats[0], batch_idx, dims.window, dims.horizon)
                x, y, edge_index = x.to(dev), y.to(dev), ei.to(dev)
                optimizer.zero_grad()
                out = temp_model(x, edge_index)
                loss = criterion(out, y)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(list(temp_model.parameters()) + list(temp_koppen.parameters()), max_norm=1.0)
                optimizer.step()
                losses.append(loss.item())
        from weatherforecast_stgcn_maml_b200 import functional as WF
        WF.check()
        assert all(np.isfinite(losses))
        after = temp_model.state_dict()
        for k in before:
            moved = not torch.equal(after[k], before[k])
            assert moved == k.startswith(("lstm.", "output_layer.")), k  # D4: no gradient reaches base_stgcn
        # train mode is stochastic, eval mode is not
        temp_model.train()
        a, b = temp_model(x, edge_index), temp_model(x, edge_index)
        assert not torch.equal(a, b)
        temp_model.eval()
        a, b = temp_model(x, edge_index), temp_model(x, edge_index)
        assert torch.equal(a, b)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
def test_module_api_dropout_matches_masked_oracle(precision):
    """HybridSTGCN_LSTM.forward / backward in train mode through autograd: masks read back per call (the module path
    snapshots (seed, pass) once per fused layer call: conv1..3 then the LSTM+head), compared with the oracle."""
    from weatherforecast_stgcn_maml_b200 import functional as WF
    from weatherforecast_stgcn_maml_b200.hybrid_model import HybridSTGCN_LSTM
    from weatherforecast_stgcn_maml_b200.model import STGCN

    WF.set_precision(precision)
    try:
        dims, ei, sd, feats = _case(G=1)
        p = 0.2
        base = STGCN(24, 256, out_channels=12, window_size=dims.window, forecast_horizon=dims.horizon, dropout_rate=p)
        hyb = HybridSTGCN_LSTM(base, lstm_hidden_size=128, lstm_num_layers=4, lstm_dropout=p, out_channels=12,
                               forecast_horizon=dims.horizon, freeze_base=False)
        hyb.load_state_dict(sd)
        hyb = hyb.cuda().train()
        st = WF._state(torch.device("cuda", torch.cuda.current_device()))
        seed, c = (int(v) for v in st.rng.tolist())
        x, y = P.window_xy(feats[0], 1, dims.window, dims.horizon)
        pred = hyb(x.cuda(), ei.cuda())
        loss = torch.nn.functional.mse_loss(pred, y.cuda())
        loss.backward()
        WF.check()
        R, N, T, L = dims.R, dims.num_nodes, dims.window, dims.lstm_hidden
        masks = {"gcn": [read_mask(seed, c + i, i, R, 256, p) for i in range(3)],
                 "lstm": [read_mask(seed, c + 3, SITE_LSTM + l, T * N, L, p).view(T, N, L).permute(1, 0, 2) for l in range(3)],
                 "head": read_mask(seed, c + 3, SITE_HEAD, N, L, p)}
        l_ref, g_ref, p_ref = P.loss_and_grads(sd, x, y, ei, dims.window, dims.horizon, 1.0, 4, masks=masks)
        assert rel_err(pred, p_ref) <= FWD_TOL
        assert abs(loss.item() - float(l_ref)) <= FWD_TOL * float(l_ref)
        named = dict(hyb.named_parameters())
        for name, ref in g_ref.items():
            assert rel_err(named[name].grad, ref) <= GRAD_TOL, name
        assert all(q.grad is None for n_, q in named.items() if n_.startswith("base_stgcn."))
    finally:
        WF.set_precision("tf32x3")


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
def test_stgcn_train_mode_dropout_after_every_conv(precision):
    """model.py:31-42: dropout after all FOUR convolutions, differentiable end to end (the GCN backward re-applies the
    forward masks): prediction, parameter gradients and dX against the oracle on the masks read back."""
    from weatherforecast_stgcn_maml_b200 import functional as WF
    from weatherforecast_stgcn_maml_b200.model import STGCN

    WF.set_precision(precision)
    try:
        dims, ei, sd, feats = _case(G=1)
        p = 0.3
        base_sd = {k[len("base_stgcn."):]: v for k, v in sd.items() if k.startswith("base_stgcn.")}
        base = STGCN(24, 256, out_channels=12, window_size=dims.window, forecast_horizon=dims.horizon, dropout_rate=p)
        base.load_state_dict(base_sd)
        base = base.cuda().train()
        st = WF._state(torch.device("cuda", torch.cuda.current_device()))
        seed, c = (int(v) for v in st.rng.tolist())
        x, y = P.window_xy(feats[0], 0, dims.window, dims.horizon)
        xs = x.cuda().requires_grad_(True)
        pred = base(xs, ei.cuda())
        torch.nn.functional.mse_loss(pred, y.cuda()).backward()
        WF.check()
        masks = [read_mask(seed, c + i, i, dims.R, 256, p) for i in range(4)]
        leaf = {k: v.clone().requires_grad_(True) for k, v in base_sd.items()}
        xr = x.clone().requires_grad_(True)
        pr = P.stgcn_forward(leaf, xr, ei, dims.window, dims.horizon, 12, masks=masks)
        gr = torch.autograd.grad(torch.nn.functional.mse_loss(pr, y), list(leaf.values()) + [xr])
        assert rel_err(pred, pr) <= FWD_TOL
        named = dict(base.named_parameters())
        for (k_, _), g in zip(leaf.items(), gr):
            assert rel_err(named[k_].grad, g) <= GRAD_TOL, k_
        assert rel_err(xs.grad, gr[-1]) <= GRAD_TOL
        base.eval()
        assert torch.equal(base(xs, ei.cuda()), base(xs, ei.cuda()))
    finally:
        WF.set_precision("tf32x3")


def test_trainers_take_dropout_from_the_model_and_default_to_the_reference():
    """inner_loop_v4 reads the three probabilities off the model (never silently off); MetaTrainer / FineTuner default
    to the reference's 0.2; a CUDA-graphed meta-step draws fresh masks on every replay."""
    from weatherforecast_stgcn_maml_b200.adapt_hybrid_v5 import FineTuner
    from weatherforecast_stgcn_maml_b200.dataset import WeatherGraphDataset
    from weatherforecast_stgcn_maml_b200.engine import REFERENCE_DROPOUT, model_dropout
    from weatherforecast_stgcn_maml_b200.hybrid_model import HybridSTGCN_LSTM
    from weatherforecast_stgcn_maml_b200.model import STGCN
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer, _runner, inner_loop_v4
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding

    assert REFERENCE_DROPOUT == (0.2, 0.2, 0.2)
    dims, ei, sd, feats = _case(G=2)
    base = STGCN(24, 256, out_channels=12, window_size=dims.window, forecast_horizon=dims.horizon, dropout_rate=0.2)
    hyb = HybridSTGCN_LSTM(base, lstm_hidden_size=128, lstm_num_layers=4, lstm_dropout=0.2, out_channels=12,
                           forecast_horizon=dims.horizon, freeze_base=False)
    hyb.load_state_dict(sd)
    assert model_dropout(hyb) == (0.2, 0.2, 0.2)
    ds = WeatherGraphDataset(feats[0], ei, dims.window, dims.horizon)
    import weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 as tm
    old = tm.INNER_EPOCHS_PER_TASK
    tm.INNER_EPOCHS_PER_TASK = 1
    try:
        m1, _ = inner_loop_v4(hyb.cuda(), KoppenEmbedding(8).cuda(), torch.utils.data.Subset(ds, [0, 1]), "cuda")
    finally:
        tm.INNER_EPOCHS_PER_TASK = old
    run = _runner(dims, 1, "cuda", (0.2, 0.2, 0.2))
    assert run.engine.stochastic and int(run.engine.rng[1].item()) == 2  # two train-mode passes consumed two mask sets
    assert m1.training

    mt = MetaTrainer(sd, [(feats[g], ei) for g in range(2)], dims, "cuda", support_rows=(0, 1), query_row=2)
    assert mt.engine.dropout == REFERENCE_DROPOUT and mt.engine.stochastic
    losses = []
    for _ in range(3):
        theta0 = mt.theta.clone()
        mt.meta_step()
        losses.append(mt.read_loss())
        mt.theta.copy_(theta0)          # same weights, same windows ...
        mt.adam.exp_avg.zero_(); mt.adam.exp_avg_sq.zero_(); mt.adam.step_count = 0
    assert len(set(losses)) == 3        # ... different masks on every graph replay
    assert int(mt.engine.rng[1].item()) >= 9
    det = MetaTrainer(sd, [(feats[g], ei) for g in range(2)], dims, "cuda", support_rows=(0, 1), query_row=2,
                      dropout=(0, 0, 0))
    a = []
    for _ in range(2):
        theta0 = det.theta.clone()
        det.meta_step()
        a.append(det.read_loss())
        det.theta.copy_(theta0)
        det.adam.exp_avg.zero_(); det.adam.exp_avg_sq.zero_(); det.adam.step_count = 0
    assert a[0] == a[1]
    ft = FineTuner(sd, feats[0], ei, dims, "cuda", region_name="India", max_samples=4, train_frac=0.5)
    assert ft.engine.dropout == REFERENCE_DROPOUT
    ft.train_epoch([0, 1])
    v1, v2 = ft.validate(), ft.validate()
    assert v1 == v2 and np.isfinite(v1)  # validation runs in eval mode (adapt_hybrid_v5.py:214)
