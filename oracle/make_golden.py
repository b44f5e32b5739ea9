"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (``python -m oracle.make_golden``): it imports
graphBuilder.py, model.py, hybrid_model.py, dataset.py, embed_utils.py and
train_hybrid_maml_v5.py straight from /root/reference over oracle/pyg_shim.py,
executes them on seeded synthetic inputs, checks oracle/ref_port.py against
them, and freezes the reference's outputs as fixtures.  /root/reference does
not exist on the GPU box; only the fixtures travel.

Dropout: the reference's training functions call ``.train()``; its three dropout
sites cannot be RNG-matched by a batched implementation, so the
reference models are *constructed* with dropout_rate=0 / lstm_dropout=0 -- the
reference's own constructor arguments, no code changed.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SAMPLE_IDX_SEED = 1234


def _import_reference():
    from oracle import build_ref

    m = build_ref.reference_modules()
    return (m["graphBuilder"], m["model"], m["hybrid_model"], m["dataset"], m["embed_utils"], m["train_hybrid_maml_v5"])


def sample_indices(numel, count=64):
    g = np.random.RandomState(SAMPLE_IDX_SEED + numel % 9973)
    return np.sort(g.choice(numel, size=min(count, numel), replace=False))


def summarize(t):
    """(l2 norm, sum, sampled values) of a tensor -- a compact fixture for big tensors."""
    a = t.detach().double().reshape(-1).numpy()
    idx = sample_indices(a.size)
    return np.array([np.sqrt((a * a).sum()), a.sum()]), a[idx].astype(np.float32)


def build_ref_model(mods, sd, cfg):
    _, model, hybrid_model, _, _, _ = mods
    base = model.STGCN(cfg["in"], cfg["hidden"], out_channels=cfg["out"], window_size=cfg["T"],
                       forecast_horizon=cfg["H"], dropout_rate=0.0)
    hyb = hybrid_model.HybridSTGCN_LSTM(base, lstm_hidden_size=cfg["L"], lstm_num_layers=cfg["layers"],
                                        lstm_dropout=0.0, out_channels=cfg["out"], forecast_horizon=cfg["H"],
                                        freeze_base=False)
    missing = hyb.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return hyb


def run_case(mods, name, cfg, nlat, nlon, k, seed, inner_steps, full_tensors):
    from oracle import ref_port as P
    from weatherforecast_stgcn_maml_b200 import synth

    graphBuilder, model, hybrid_model, dataset, embed_utils, train = mods
    T, H = cfg["T"], cfg["H"]
    lats, lons = synth.region_grid(nlat, nlon)
    with contextlib.redirect_stdout(io.StringIO()):
        edge_index, n, pos = graphBuilder.build_spatial_graph(synth.GridCoords(lats, lons), k_neighbors=k)
    assert torch.equal(edge_index, P.knn_edges_ckdtree(lats, lons, k))
    sd = synth.init_v5_state_dict(seed, gcn_bias_scale=0.05, in_channels=cfg["in"], hidden=cfg["hidden"],
                                  lstm_hidden=cfg["L"], lstm_layers=cfg["layers"], out_channels=cfg["out"],
                                  horizon=H)
    feats = synth.synth_features(T + H + 8, n, seed + 1, synth.koppen_table(seed)[3])
    ds = dataset.WeatherGraphDataset(feats, edge_index, window_size=T, forecast_horizon=H)
    assert len(ds) == P.num_windows(feats, T, H)
    out = {"edge_index": edge_index.numpy().astype(np.int32), "k": k, "nlat": nlat, "nlon": nlon,
           "seed": seed, "cfg": np.array([cfg[x] for x in ("in", "hidden", "L", "layers", "out", "T", "H")])}

    hyb = build_ref_model(mods, sd, cfg)
    assert sum(p.numel() for p in hyb.parameters()) == sum(v.numel() for v in sd.values())
    out["state_dict_keys"] = np.array(list(hyb.state_dict().keys()))
    assert list(hyb.state_dict().keys()) == list(sd.keys())

    # ---- forward / loss / backward on window 0, eval mode (hybrid_model.py:80-117)
    hyb.train()  # dropout p = 0 everywhere
    d0 = ds[0]
    x_p, y_p = P.window_xy(feats, 0, T, H)
    assert torch.equal(d0.x, x_p) and torch.equal(d0.y, y_p)
    pred = hyb(d0.x, d0.edge_index)
    loss = torch.nn.MSELoss()(pred, d0.y)
    loss.backward()
    ref_grads = {k_: p.grad for k_, p in hyb.named_parameters()}
    assert all(ref_grads[k_] is None for k_ in sd if k_.startswith("base_stgcn."))
    p_loss, p_grads, p_pred = P.loss_and_grads(sd, x_p, y_p, edge_index, T, H, 1.0, cfg["layers"])
    err = (p_pred - pred).abs().max().item() / pred.abs().max().item()
    assert err < 2e-6, f"port forward vs reference: {err}"
    for k_ in p_grads:
        e = (p_grads[k_] - ref_grads[k_]).abs().max().item() / (ref_grads[k_].abs().max().item() + 1e-30)
        assert e < 2e-5, f"port grad {k_}: {e}"
    out["pred"] = pred.detach().numpy()
    out["loss"] = np.float64(loss.item())

    # base features (hybrid_model.py:60-78)
    bf = hyb.extract_base_features(d0.x, d0.edge_index)
    pf = P.gcn_stack(sd, x_p, edge_index)
    assert (bf - pf).abs().max().item() <= 1e-6 * bf.abs().max().item()
    out["base_features_summary"], out["base_features_samples"] = summarize(bf)
    if full_tensors:
        out["base_features"] = bf.numpy()

    for k_, g in ref_grads.items():
        if g is None:
            continue
        s, v = summarize(g)
        out[f"grad_summary/{k_}"], out[f"grad_samples/{k_}"] = s, v
        if full_tensors:
            out[f"grad/{k_}"] = g.numpy()

    # ---- STGCN.forward fwd+bwd (model.py:30-52): the only differentiable GCN use (D4)
    base_sd = {k_[len("base_stgcn."):]: v for k_, v in sd.items() if k_.startswith("base_stgcn.")}
    base = model.STGCN(cfg["in"], cfg["hidden"], out_channels=cfg["out"], window_size=T,
                       forecast_horizon=H, dropout_rate=0.0)
    base.load_state_dict(base_sd)
    base.train()
    xs = d0.x.clone().requires_grad_(True)
    sp = base(xs, d0.edge_index)
    sl = torch.nn.MSELoss()(sp, d0.y)
    sl.backward()
    leaf = {k_: v.clone().requires_grad_(True) for k_, v in base_sd.items()}
    xs2 = d0.x.clone().requires_grad_(True)
    pp = P.stgcn_forward(leaf, xs2, edge_index, T, H, cfg["out"])
    pl = torch.nn.functional.mse_loss(pp, d0.y)
    pg = torch.autograd.grad(pl, list(leaf.values()) + [xs2])
    assert (pp - sp).abs().max().item() <= 2e-6 * sp.abs().max().item()
    for (k_, _), g in zip(list(leaf.items()), pg):
        rg = dict(base.named_parameters())[k_].grad
        assert (g - rg).abs().max().item() <= 2e-5 * (rg.abs().max().item() + 1e-30), k_
    assert (pg[-1] - xs.grad).abs().max().item() <= 2e-5 * xs.grad.abs().max().item()
    out["stgcn_pred"] = sp.detach().numpy()
    out["stgcn_loss"] = np.float64(sl.item())
    for k_, p in base.named_parameters():
        s, v = summarize(p.grad)
        out[f"stgcn_grad_summary/{k_}"], out[f"stgcn_grad_samples/{k_}"] = s, v
        if full_tensors:
            out[f"stgcn_grad/{k_}"] = p.grad.numpy()
    out["stgcn_dx_summary"], out["stgcn_dx_samples"] = summarize(xs.grad)
    if full_tensors:
        out["stgcn_dx"] = xs.grad.numpy()

    # ---- reference inner_loop_v4 + query backward (train_hybrid_maml_v5.py:110-141,162-169)
    from torch.utils.data import Subset

    hyb2 = build_ref_model(mods, sd, cfg)
    kop = embed_utils.KoppenEmbedding(8)
    support = Subset(ds, list(range(inner_steps)))
    query = Subset(ds, [inner_steps])
    train.INNER_EPOCHS_PER_TASK = 1  # module constant; the loop body is untouched
    adapted, _ = train.inner_loop_v4(hyb2, kop, support, "cpu")
    adapted.train()
    qb = next(iter(train.DataLoader(query, batch_size=1, shuffle=False)))
    # The copy still carries the last inner step's clipped grads (zero_grad runs at the START of
    # each inner step, train_hybrid_maml_v5.py:130), so the reference's query backward (:169)
    # ACCUMULATES onto them.  Both readings are frozen: "literal" (stale + query) and "fomaml"
    # (query only, the meta-gradient this build defines).
    stale = {k_: p.grad.clone() for k_, p in adapted.named_parameters() if p.grad is not None}
    qout = adapted(qb.x, qb.edge_index)
    qloss = torch.nn.MSELoss()(qout, qb.y) / train.GRAD_ACCUMULATION_STEPS
    qloss.backward()
    literal = {k_: p.grad.clone() for k_, p in adapted.named_parameters() if p.grad is not None}
    assert all(p.grad is None for p in hyb2.parameters())  # nothing reaches the meta-params
    adapted.zero_grad()
    qloss2 = torch.nn.MSELoss()(adapted(qb.x, qb.edge_index), qb.y) / train.GRAD_ACCUMULATION_STEPS
    qloss2.backward()
    for k_, g in literal.items():
        pure = dict(adapted.named_parameters())[k_].grad
        assert (g - (stale[k_] + pure)).abs().max().item() <= 1e-6 * (g.abs().max().item() + 1e-30)
        s, v = summarize(g)
        out[f"literal_summary/{k_}"], out[f"literal_samples/{k_}"] = s, v
        s, v = summarize(stale[k_])
        out[f"stale_summary/{k_}"], out[f"stale_samples/{k_}"] = s, v
    p_q, p_qg, p_fast = P.fomaml_task(sd, feats, edge_index, list(range(inner_steps)), inner_steps,
                                      train.GRAD_ACCUMULATION_STEPS, window=T, horizon=H, lr=train.INNER_LR,
                                      lstm_layers=cfg["layers"])
    ad_sd = adapted.state_dict()
    for k_ in P.trainable(sd):
        e = (p_fast[k_] - ad_sd[k_]).abs().max().item() / ad_sd[k_].abs().max().item()
        assert e < 2e-5, f"port adapted {k_}: {e}"
        g = dict(adapted.named_parameters())[k_].grad
        e = (p_qg[k_] - g).abs().max().item() / (g.abs().max().item() + 1e-30)
        assert e < 1e-4, f"port fomaml grad {k_}: {e}"
        s, v = summarize(ad_sd[k_])
        out[f"adapted_summary/{k_}"], out[f"adapted_samples/{k_}"] = s, v
        s, v = summarize(g)
        out[f"fomaml_summary/{k_}"], out[f"fomaml_samples/{k_}"] = s, v
        if full_tensors:
            out[f"adapted/{k_}"] = ad_sd[k_].numpy()
            out[f"fomaml/{k_}"] = g.numpy()
    for k_ in sd:
        if k_.startswith("base_stgcn."):
            assert torch.equal(ad_sd[k_], sd[k_])  # SGD skips grad=None params
    out["inner_steps"] = inner_steps
    out["accum"] = train.GRAD_ACCUMULATION_STEPS
    out["inner_lr"] = train.INNER_LR
    out["query_loss_scaled"] = np.float64(qloss.item())
    assert abs(float(p_q) - qloss.item()) <= 1e-5 * abs(qloss.item())

    if full_tensors:
        for k_, v in sd.items():
            out[f"sd/{k_}"] = v.numpy()
        out["features"] = feats.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"[golden] {name}: loss={loss.item():.6f} stgcn_loss={sl.item():.6f} qloss/accum={qloss.item():.6f}")


def knn_cases(mods):
    from oracle import ref_port as P
    from weatherforecast_stgcn_maml_b200 import synth

    graphBuilder = mods[0]
    out = {}
    for (nlat, nlon, k) in [(21, 21, 4), (21, 21, 8), (5, 7, 4), (3, 3, 2), (121, 121, 8), (40, 30, 4)]:
        lats, lons = synth.region_grid(nlat, nlon)
        with contextlib.redirect_stdout(io.StringIO()):
            ei, n, pos = graphBuilder.build_spatial_graph(synth.GridCoords(lats, lons), k_neighbors=k)
        assert n == nlat * nlon and ei.shape == (2, n * k) and ei.dtype == torch.int64
        out[f"{nlat}x{nlon}_k{k}"] = ei.numpy().astype(np.int32)
        if n <= 2000:
            can = P.knn_edges_canonical(lats, lons, k)
            d_ref = np.sort(P.knn_sq_distances(lats, lons, ei).reshape(n, k), axis=1)
            d_can = np.sort(P.knn_sq_distances(lats, lons, can).reshape(n, k), axis=1)
            assert np.array_equal(d_ref, d_can), "distance multisets must agree everywhere"
    np.savez_compressed(os.path.join(OUT, "knn_ckdtree.npz"), **out)
    print("[golden] knn_ckdtree:", {k_: v.shape for k_, v in out.items()})


class _Var:
    def __init__(self, values):
        self.values = values


def synth_raw_weather(seed, time, nlat, nlon, nan_frac=0.01, all_nan_var=None):
    """Raw (un-normalised) ERA5-like fields f32 [time, lat, lon, 12] with physical offsets and scales, NaNs sprinkled in
    three variables (and optionally one variable entirely NaN), plus day-of-year / hour-of-day of hourly steps."""
    g = np.random.RandomState(seed)
    loc = np.array([1.5, -0.7, 288.0, 281.0, 98000.0, 2e-4, 2.5, -1.1, -6.0e4, 0.35, 0.45, -8e-5], dtype=np.float64)
    scale = np.array([3.0, 3.0, 9.0, 8.0, 2500.0, 5e-4, 4.5, 4.5, 3.0e4, 0.3, 0.3, 6e-5], dtype=np.float64)
    w = (loc + scale * g.standard_normal((time, nlat, nlon, 12))).astype(np.float32)
    for v in (2, 5, 9):
        w[..., v][g.random_sample((time, nlat, nlon)) < nan_frac] = np.nan
    if all_nan_var is not None:
        w[..., all_nan_var] = np.nan
    hours = 17 + np.arange(time)
    return w, 200 + hours // 24, (hours % 24).astype(np.float64)


def feature_cases():
    """prepare_model_input of the UNMODIFIED reference (featurePreprocessor.py:66-182) on a dict-like stand-in for the
    xarray Dataset; embed_utils.add_time_embeddings needs a real Dataset, so the time features come from the restated
    formula (ref_port.time_features, embed_utils.py:12-26).  Freezes inputs, outputs and statistics."""
    from oracle import pyg_shim, ref_port

    pyg_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import featurePreprocessor as fp
        import embed_utils
    torch.manual_seed(42)
    kop = embed_utils.KoppenEmbedding(8)
    out = {"koppen_weight": kop.embedding.weight.detach().numpy().copy()}
    given = {"mean": [float(x) for x in np.linspace(-2.0, 3.0, 12)], "std": [float(x) for x in np.linspace(0.5, 4.0, 12)]}
    for name, kw, code, stats in (("nan", dict(), 7, None), ("allnan", dict(all_nan_var=4), 12, None),
                                  ("clean", dict(nan_frac=0.0), 3, None), ("given", dict(), 21, given),
                                  ("raw", dict(), 5, "raw")):
        w, doy, tod = synth_raw_weather(11 + code, 40, 5, 7, **kw)
        tf = ref_port.time_features(doy, tod)
        ds = {v: _Var(w[..., i].copy()) for i, v in enumerate(fp.WEATHER_VARS)}
        ds.update({v: _Var(tf[:, i]) for i, v in enumerate(fp.TIME_VARS)})
        normalize = stats != "raw"
        with contextlib.redirect_stdout(io.StringIO()):
            feats, st = fp.prepare_model_input(ds, code, kop, normalize=normalize, stats=stats if normalize else None)
        mine, st2 = ref_port.prepare_features(w.copy(), tf, kop(torch.tensor([code])), normalize=normalize,
                                              stats=stats if normalize else None)
        assert torch.equal(feats, mine), name
        if normalize:
            assert np.array_equal(np.asarray(st["mean"]), np.asarray(st2["mean"])), name
        out[f"{name}_weather"], out[f"{name}_doy"], out[f"{name}_tod"] = w, doy, tod
        out[f"{name}_code"] = np.array(code)
        out[f"{name}_features"] = feats.detach().numpy()
        if normalize:
            out[f"{name}_mean"], out[f"{name}_std"] = np.asarray(st["mean"]), np.asarray(st["std"])
    np.savez_compressed(os.path.join(OUT, "features_prepare.npz"), **out)
    print("features_prepare.npz:", {k: v.shape for k, v in out.items() if k.endswith("_features")})


def inner90_case(mods, name="hybrid_v5_k8_inner90", nlat=21, nlon=21, k=8, seed=43):
    """The reference's REAL inner-loop shape (train_hybrid_maml_v5.py:124-127): INNER_EPOCHS_PER_TASK = 6 passes over the
    first 15 support windows = 90 sequential SGD steps with the unmodified ``inner_loop_v4``, then the query backward
    (:162-169).  Freezes the adapted weights and the first-order meta-gradient so that the drift of a reduced-precision
    path over the full loop is pinned (about 6 minutes of CPU)."""
    from torch.utils.data import Subset

    from oracle import ref_port as P
    from weatherforecast_stgcn_maml_b200 import synth

    graphBuilder, model, hybrid_model, dataset, embed_utils, train = mods
    cfg = dict(T=24, H=8, hidden=256, L=128, layers=4, out=12)
    cfg["in"] = 24
    T, H = cfg["T"], cfg["H"]
    lats, lons = synth.region_grid(nlat, nlon)
    with contextlib.redirect_stdout(io.StringIO()):
        edge_index, n, pos = graphBuilder.build_spatial_graph(synth.GridCoords(lats, lons), k_neighbors=k)
    sd = synth.init_v5_state_dict(seed, gcn_bias_scale=0.05, in_channels=24, hidden=256, lstm_hidden=128, lstm_layers=4,
                                  out_channels=12, horizon=H)
    nsup = train_per_epoch = 15
    feats = synth.synth_features(T + H + nsup + 4, n, seed + 1, synth.koppen_table(seed)[3])
    ds = dataset.WeatherGraphDataset(feats, edge_index, window_size=T, forecast_horizon=H)
    hyb = build_ref_model(mods, sd, cfg)
    kop = embed_utils.KoppenEmbedding(8)
    train.INNER_EPOCHS_PER_TASK = 6  # the reference's own value (:24); set explicitly because run_case lowers it
    support = Subset(ds, list(range(nsup + 2)))  # more than 15 windows: the loop's own `break` at 15 (:126) cuts it
    query = Subset(ds, [nsup])
    adapted, _ = train.inner_loop_v4(hyb, kop, support, "cpu")
    adapted.train()
    qb = next(iter(train.DataLoader(query, batch_size=1, shuffle=False)))
    adapted.zero_grad()
    qloss = torch.nn.MSELoss()(adapted(qb.x, qb.edge_index), qb.y) / train.GRAD_ACCUMULATION_STEPS
    qloss.backward()
    steps = P.reference_support_schedule_indices(nsup + 2, 6, train_per_epoch)
    assert len(steps) == 90
    p_q, p_qg, p_fast = P.fomaml_task(sd, feats, edge_index, steps, nsup, train.GRAD_ACCUMULATION_STEPS, window=T,
                                      horizon=H, lr=train.INNER_LR, lstm_layers=4)
    out = {"edge_index": edge_index.numpy().astype(np.int32), "k": k, "nlat": nlat, "nlon": nlon, "seed": seed,
           "cfg": np.array([cfg[x] for x in ("in", "hidden", "L", "layers", "out", "T", "H")]),
           "inner_steps": 90, "support_windows": nsup, "query_window": nsup, "accum": train.GRAD_ACCUMULATION_STEPS,
           "inner_lr": train.INNER_LR, "query_loss_scaled": np.float64(qloss.item()), "feature_rows": feats.shape[0]}
    ad_sd = adapted.state_dict()
    for k_ in P.trainable(sd):
        e = (p_fast[k_] - ad_sd[k_]).abs().max().item() / ad_sd[k_].abs().max().item()
        assert e < 1e-4, f"port adapted (90 steps) {k_}: {e}"
        g = dict(adapted.named_parameters())[k_].grad
        e = (p_qg[k_] - g).abs().max().item() / (g.abs().max().item() + 1e-30)
        assert e < 1e-3, f"port fomaml grad (90 steps) {k_}: {e}"
        out[f"adapted_summary/{k_}"], out[f"adapted_samples/{k_}"] = summarize(ad_sd[k_])
        out[f"fomaml_summary/{k_}"], out[f"fomaml_samples/{k_}"] = summarize(g)
        out[f"delta_summary/{k_}"], out[f"delta_samples/{k_}"] = summarize(ad_sd[k_] - sd[k_])  # what 90 steps moved
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"[golden] {name}: 90 inner steps, qloss/accum={qloss.item():.6f}")


def _replay_masks(shapes, p):
    """The keep-masks ATen's CPU dropout draws, in call order: ``at::dropout`` = ``empty_like(x).bernoulli_(1 - p)
    .div_(1 - p)`` on the default generator -- used both by nn.Dropout / F.dropout and between nn.LSTM layers."""
    return [torch.empty(s).bernoulli_(1 - p).div_(1 - p) for s in shapes]


def dropout_case(mods, name="hybrid_small_dropout", nlat=3, nlon=4, k=4, seed=7, p=0.2, rng_seed=2024):
    """TRAIN-MODE reference (dropout_rate = lstm_dropout = p > 0, ``.train()``) with torch's generator seeded: the masks
    the unmodified reference draws are replayed from the same seed in the same order -- three GCN sites
    (hybrid_model.py:67-73), per node the nn.LSTM inter-layer sites (:47, one call per node, :98), the head site (:108)
    -- and ``ref_port.hybrid_forward(masks=...)`` must reproduce the reference's predictions, loss and gradients.  This
    pins the PLACEMENT and SCALING of every dropout site of the restatement (and so of the CUDA path, which is tested
    against the restatement on masks read back from the kernels) against the reference itself."""
    from oracle import ref_port as P
    from weatherforecast_stgcn_maml_b200 import synth

    graphBuilder, model, hybrid_model, dataset, embed_utils, train = mods
    cfg = dict(T=6, H=2, hidden=32, L=32, layers=3, out=12)
    cfg["in"] = 24
    T, H, L, Ls = cfg["T"], cfg["H"], cfg["L"], cfg["layers"]
    lats, lons = synth.region_grid(nlat, nlon)
    with contextlib.redirect_stdout(io.StringIO()):
        edge_index, n, pos = graphBuilder.build_spatial_graph(synth.GridCoords(lats, lons), k_neighbors=k)
    sd = synth.init_v5_state_dict(seed, gcn_bias_scale=0.05, in_channels=24, hidden=cfg["hidden"], lstm_hidden=L,
                                  lstm_layers=Ls, out_channels=12, horizon=H)
    feats = synth.synth_features(T + H + 8, n, seed + 1, synth.koppen_table(seed)[3])
    x, y = P.window_xy(feats, 0, T, H)
    base = model.STGCN(24, cfg["hidden"], out_channels=12, window_size=T, forecast_horizon=H, dropout_rate=p)
    hyb = hybrid_model.HybridSTGCN_LSTM(base, lstm_hidden_size=L, lstm_num_layers=Ls, lstm_dropout=p, out_channels=12,
                                        forecast_horizon=H, freeze_base=False)
    hyb.load_state_dict(sd)
    hyb.train()
    torch.set_num_threads(1)  # the replay must consume the generator exactly as the reference run does
    torch.manual_seed(rng_seed)
    pred = hyb(x, edge_index)
    loss = torch.nn.MSELoss()(pred, y)
    loss.backward()
    torch.manual_seed(rng_seed)
    R = T * n
    gcn = _replay_masks([(R, cfg["hidden"])] * 3, p)
    lstm = [torch.empty(n, T, L) for _ in range(Ls - 1)]
    for node in range(n):
        for l, m in enumerate(_replay_masks([(1, T, L)] * (Ls - 1), p)):  # batch_first input [1, T, L] of one node
            lstm[l][node] = m[0]
    head = _replay_masks([(n, L)], p)[0]
    masks = {"gcn": gcn, "lstm": lstm, "head": head}
    l_p, g_p, p_p = P.loss_and_grads(sd, x, y, edge_index, T, H, 1.0, Ls, masks=masks)
    err = (p_p - pred).abs().max().item() / pred.abs().max().item()
    assert err < 2e-6, f"masked port vs train-mode reference: forward {err} (mask replay out of step?)"
    out = {"edge_index": edge_index.numpy().astype(np.int32), "k": k, "nlat": nlat, "nlon": nlon, "seed": seed, "p": p,
           "cfg": np.array([cfg[x_] for x_ in ("in", "hidden", "L", "layers", "out", "T", "H")]),
           "pred": pred.detach().numpy(), "loss": np.float64(loss.item()), "features": feats.numpy(),
           "mask_head": head.numpy()}
    for i, m in enumerate(gcn):
        out[f"mask_gcn{i}"] = m.numpy()
    for l, m in enumerate(lstm):
        out[f"mask_lstm{l}"] = m.numpy()
    for k_, v in sd.items():
        out[f"sd/{k_}"] = v.numpy()
    for k_, q in hyb.named_parameters():
        if k_.startswith("base_stgcn."):
            assert q.grad is None
            continue
        e = (g_p[k_] - q.grad).abs().max().item() / (q.grad.abs().max().item() + 1e-30)
        assert e < 2e-5, f"masked port grad {k_}: {e}"
        out[f"grad/{k_}"] = q.grad.numpy()
    # STGCN.forward in train mode: dropout after all four convolutions (model.py:31-42), differentiable
    base_sd = {k_[len("base_stgcn."):]: v for k_, v in sd.items() if k_.startswith("base_stgcn.")}
    b2 = model.STGCN(24, cfg["hidden"], out_channels=12, window_size=T, forecast_horizon=H, dropout_rate=p)
    b2.load_state_dict(base_sd)
    b2.train()
    torch.manual_seed(rng_seed + 1)
    xs = x.clone().requires_grad_(True)
    sp = b2(xs, edge_index)
    sl = torch.nn.MSELoss()(sp, y)
    sl.backward()
    torch.manual_seed(rng_seed + 1)
    m4 = _replay_masks([(R, cfg["hidden"])] * 4, p)
    leaf = {k_: v.clone().requires_grad_(True) for k_, v in base_sd.items()}
    xs2 = x.clone().requires_grad_(True)
    pp = P.stgcn_forward(leaf, xs2, edge_index, T, H, 12, masks=m4)
    pg = torch.autograd.grad(torch.nn.functional.mse_loss(pp, y), list(leaf.values()) + [xs2])
    assert (pp - sp).abs().max().item() <= 2e-6 * sp.abs().max().item(), "masked STGCN port vs train-mode reference"
    assert (pg[-1] - xs.grad).abs().max().item() <= 2e-5 * xs.grad.abs().max().item()
    out["stgcn_pred"], out["stgcn_loss"], out["stgcn_dx"] = sp.detach().numpy(), np.float64(sl.item()), xs.grad.numpy()
    for i, m in enumerate(m4):
        out[f"stgcn_mask{i}"] = m.numpy()
    for k_, q in b2.named_parameters():
        out[f"stgcn_grad/{k_}"] = q.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"[golden] {name}: train-mode loss={loss.item():.6f} stgcn_loss={sl.item():.6f} (p={p})")


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "features":  # only this fixture (the model cases take minutes)
        feature_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] in ("inner90", "dropout"):
        torch.manual_seed(42)
        np.random.seed(42)
        mods = _import_reference()
        if sys.argv[1] == "inner90":
            torch.set_num_threads(os.cpu_count() or 1)
            inner90_case(mods)
        else:
            dropout_case(mods)
        return
    torch.manual_seed(42)
    np.random.seed(42)
    torch.set_num_threads(os.cpu_count() or 1)
    mods = _import_reference()
    knn_cases(mods)
    small = dict(T=6, H=2, hidden=32, L=32, layers=2, out=12)
    small["in"] = 24
    run_case(mods, "hybrid_small", small, nlat=3, nlon=4, k=4, seed=7, inner_steps=3, full_tensors=True)
    full = dict(T=24, H=8, hidden=256, L=128, layers=4, out=12)
    full["in"] = 24
    run_case(mods, "hybrid_v5_k4", full, nlat=21, nlon=21, k=4, seed=42, inner_steps=3, full_tensors=False)
    run_case(mods, "hybrid_v5_k8", full, nlat=21, nlon=21, k=8, seed=43, inner_steps=3, full_tensors=False)
    feature_cases()
    dropout_case(mods)
    inner90_case(mods)


if __name__ == "__main__":
    main()
