"""GPU parity of the training loops through the C ABI: fused clip+SGD / Adam(W), inner loop,
FOMAML meta-gradient, meta_update_v4 (both readings of SURVEY.md D5), MetaTrainer, fine-tuning."""
import copy

import numpy as np
import pytest
import torch

from conftest import check_summary, golden_case, rel_err
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-4, 1e-3


def _small_cfg():
    z, cfg, sd, feats, ei = golden_case("hybrid_small")
    from weatherforecast_stgcn_maml_b200.engine import V5Dims

    dims = V5Dims(num_nodes=cfg["nlat"] * cfg["nlon"], window=cfg["T"], horizon=cfg["H"], in_channels=cfg["cin"],
                  hidden=cfg["hidden"], lstm_hidden=cfg["L"], lstm_layers=cfg["layers"], out_channels=cfg["out"])
    return z, cfg, sd, feats, ei, dims


def _hybrid(cfg, sd):
    from weatherforecast_stgcn_maml_b200.hybrid_model import HybridSTGCN_LSTM
    from weatherforecast_stgcn_maml_b200.model import STGCN

    base = STGCN(cfg["cin"], cfg["hidden"], out_channels=cfg["out"], window_size=cfg["T"],
                 forecast_horizon=cfg["H"], dropout_rate=0.0)
    hyb = HybridSTGCN_LSTM(base, lstm_hidden_size=cfg["L"], lstm_num_layers=cfg["layers"], lstm_dropout=0.0,
                           out_channels=cfg["out"], forecast_horizon=cfg["H"], freeze_base=False)
    hyb.load_state_dict(sd)
    return hyb.cuda()


@pytest.mark.parametrize("scale", [0.01, 30.0])
def test_clip_sgd_matches_torch(scale):
    from weatherforecast_stgcn_maml_b200 import _lib

    torch.manual_seed(0)
    G, Pn = 3, 4096 + 64
    theta = torch.randn(G, Pn)
    grad = torch.randn(G, Pn) * scale
    ref = theta.clone()
    norms = []
    for g in range(G):
        p = torch.nn.Parameter(ref[g].clone())
        p.grad = grad[g].clone()
        norms.append(torch.nn.utils.clip_grad_norm_([p], 1.0))
        torch.optim.SGD([p], lr=0.01).step()
        ref[g] = p.detach()
    th, gr = theta.cuda(), grad.cuda()
    nr = torch.zeros(G, device="cuda")
    nb = _lib.query("wf_optim_workspace_bytes", G)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    _lib.call("wf_clip_sgd_step", _lib.ptr(th), Pn, _lib.ptr(gr), Pn, Pn, G, 0.01, 1.0, _lib.ptr(nr), _lib.ptr(ws), nb,
              _lib.stream_ptr())
    assert rel_err(th, ref) <= 1e-6
    assert rel_err(nr, torch.stack(norms)) <= 1e-5


@pytest.mark.parametrize("decoupled", [True, False])
def test_clip_adam_matches_torch(decoupled):
    from weatherforecast_stgcn_maml_b200.engine import AdamState

    torch.manual_seed(1)
    Pn = 10000
    p = torch.nn.Parameter(torch.randn(Pn))
    opt = (torch.optim.AdamW([p], lr=1e-3, weight_decay=1e-4) if decoupled
           else torch.optim.Adam([p], lr=6e-4, weight_decay=1e-4))
    theta = p.detach().clone().cuda()
    st = AdamState(Pn, "cuda", 1e-3 if decoupled else 6e-4, weight_decay=1e-4, decoupled=decoupled)
    for step in range(5):
        g = torch.randn(Pn) * (5.0 if step % 2 else 0.001)
        p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([p], 1.0)
        opt.step()
        st.step(theta, g.cuda(), max_norm=1.0)
        if step == 2:
            for grp in opt.param_groups:
                grp["lr"] = 3e-4
            st.lr = 3e-4
    assert rel_err(theta, p.detach()) <= 2e-6


@pytest.mark.parametrize("name", ["hybrid_v5_k4", "hybrid_v5_k8"])
def test_inner_loop_and_fomaml_vs_reference_fixture(name):
    """3 inner SGD steps + query backward on the v5 model (tensor-core path: the one the benchmark runs, k = 8 being the
    benchmark's k), against the reference's own run."""
    from weatherforecast_stgcn_maml_b200.engine import V5Dims, unflatten_trainable
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer

    z, cfg, sd, feats, ei = golden_case(name)
    dims = V5Dims(num_nodes=441)
    steps, accum = int(z["inner_steps"]), int(z["accum"])
    mt = MetaTrainer(sd, [(feats, ei)], dims, "cuda", dropout=(0, 0, 0), support_rows=tuple(range(steps)), query_row=steps,
                     inner_lr=float(z["inner_lr"]), accum=accum, use_cuda_graph=False)
    loss = mt.meta_step()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(z["query_loss_scaled"])) <= FWD_TOL * float(z["query_loss_scaled"])
    fast = unflatten_trainable(mt.fast[0].cpu(), dims)
    mg = unflatten_trainable(mt.meta_gradient().cpu(), dims)
    assert mt.engine.seq, "this test must run the persistent tensor-core path"
    mt.check()
    for k in fast:
        check_summary(fast[k], z[f"adapted_summary/{k}"], z[f"adapted_samples/{k}"], GRAD_TOL, "adapted " + k)
        check_summary(mg[k], z[f"fomaml_summary/{k}"], z[f"fomaml_samples/{k}"], GRAD_TOL, "fomaml " + k)


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
def test_reference_inner_loop_shape_90_steps_vs_reference_fixture(precision):
    """The reference's real inner loop: 6 epochs x the first 15 support windows = 90 SEQUENTIAL clip+SGD steps
    (train_hybrid_maml_v5.py:124-127), then the query backward (:162-169) -- against the unmodified reference's own
    90-step run.  Shows that the 16-bit hi/lo operand splits do not drift: adapted weights, the weight CHANGE the loop
    produced, and the first-order meta-gradient all stay within 1e-3."""
    from conftest import load_golden
    from weatherforecast_stgcn_maml_b200.engine import (HybridEngine, V5Dims, flatten_trainable,
                                                        gcn_weights_from_state_dict, unflatten_trainable)
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph

    z = load_golden("hybrid_v5_k8_inner90")
    seed, T, H = int(z["seed"]), 24, 8
    dims = V5Dims(num_nodes=441)
    sd = synth.init_v5_state_dict(seed, gcn_bias_scale=0.05)
    feats = synth.synth_features(int(z["feature_rows"]), 441, seed + 1, synth.koppen_table(seed)[3])
    ei = torch.from_numpy(z["edge_index"].astype(np.int64))
    steps = P.reference_support_schedule_indices(int(z["support_windows"]) + 2, 6, 15)
    assert len(steps) == int(z["inner_steps"]) == 90
    dev = "cuda"
    eng = HybridEngine(dims, 1, 1, dev, precision=precision)
    assert eng.seq == (precision == "tf32x3")
    graph = RegionGraph(ei, dims.R, dev)
    fd = feats.to(dev)
    per = 441 * 24
    gw = gcn_weights_from_state_dict(sd, dev)
    fast = flatten_trainable(sd, dims, dev).clone().unsqueeze(0)
    theta0 = fast.clone()
    xo = torch.zeros(1, dtype=torch.long, device=dev)
    to = torch.zeros(1, dtype=torch.long, device=dev)

    def one(idx, scale):
        xo.fill_(idx * per)
        to.fill_((idx + T + 1) * per)
        return eng.forward_backward(fd, 24, 0, xo, gw, graph, fast, eng.P, feat=fd, tgt_off=to, feat_ld=24, grad_scale=scale)

    for idx in steps:
        one(idx, 1.0)
        eng.sgd_step(fast, float(z["inner_lr"]), 1.0)
    loss, grads = one(int(z["query_window"]), 1.0 / int(z["accum"]))
    torch.cuda.synchronize()
    eng.check()
    ql = loss[0].item() / int(z["accum"])
    assert abs(ql - float(z["query_loss_scaled"])) <= 1e-3 * float(z["query_loss_scaled"])
    got = unflatten_trainable(fast[0].cpu(), dims)
    delta = unflatten_trainable((fast[0] - theta0[0]).cpu(), dims)
    mg = unflatten_trainable(grads[0].cpu(), dims)
    for k in got:
        check_summary(got[k], z[f"adapted_summary/{k}"], z[f"adapted_samples/{k}"], GRAD_TOL, "adapted(90) " + k)
        check_summary(mg[k], z[f"fomaml_summary/{k}"], z[f"fomaml_samples/{k}"], GRAD_TOL, "fomaml(90) " + k)
        # the weights barely move relative to their size, so also pin what the 90 steps CHANGED (a much harder test;
        # the fixture's own fp32 summation order is only repeatable to ~1e-3 of this small quantity after 90 steps)
        check_summary(delta[k], z[f"delta_summary/{k}"], z[f"delta_samples/{k}"], 5e-3, "delta(90) " + k)


def test_meta_trainer_two_tasks_vs_oracle_with_and_without_graph():
    """Two tasks, different graphs; FOMAML gradient sum + fused AdamW vs oracle + torch.optim.AdamW."""
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer

    z, cfg, sd, feats, ei, dims = _small_cfg()
    lats, lons = synth.region_grid(cfg["nlat"], cfg["nlon"])
    ei2 = P.knn_edges_canonical(lats, lons, cfg["k"])
    feats2 = synth.synth_features(feats.shape[0], feats.shape[1], 77)
    tasks = [(feats, ei), (feats2, ei2)]
    kw = dict(window=cfg["T"], horizon=cfg["H"], lr=0.01, lstm_layers=cfg["layers"])
    names = P.trainable(sd)
    # oracle: two meta-steps, accum = 2, AdamW(lr 1e-3, wd 1e-4), clip 1.0
    cur = {k: v.clone() for k, v in sd.items()}
    params = [torch.nn.Parameter(cur[k].clone()) for k in names]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-4)
    ref_losses = []
    for it in range(2):
        tot, gsum = 0.0, None
        for f, e in tasks:
            l, g, _ = P.fomaml_task(cur, f, e, [0, 1, 2], 3, 2, **kw)
            tot += float(l)
            gsum = g if gsum is None else {k: gsum[k] + g[k] for k in g}
        for p, k in zip(params, names):
            p.grad = gsum[k].clone()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        for p, k in zip(params, names):
            cur[k] = p.detach().clone()
        ref_losses.append(tot)
    # eager; one graph per meta-step with the all-reduce + AdamW inside it (default) or enqueued after it; host-resident
    # features (two graphs, double-buffered uploads: capture must not run the outer update)
    for use_graph, fused, host in ((False, None, False), (True, True, False), (True, False, False), (True, True, True)):
        mt = MetaTrainer(sd, tasks, dims, "cuda", dropout=(0, 0, 0), support_rows=(0, 1, 2), query_row=3, accum=2,
                         use_cuda_graph=use_graph, fused_update=fused, host_staging=host)
        assert mt.fused_update == bool(use_graph and fused)
        got = [mt.meta_step().item() for _ in range(2)]
        torch.cuda.synchronize()
        assert mt.adam.step_count == 2
        out = mt.state_dict()
        assert np.allclose(got, ref_losses, rtol=FWD_TOL), (use_graph, fused, host, got, ref_losses)
        for k in names:
            assert rel_err(out[k], cur[k]) <= 1e-4, (use_graph, fused, host, k)
        for k in sd:
            if k.startswith("base_stgcn."):
                assert torch.equal(out[k], sd[k])


def _ref_tasks(cfg, feats_list, ei_list, n_support, dataset_mod):
    from torch.utils.data import Subset

    tasks = []
    for f, e in zip(feats_list, ei_list):
        ds = dataset_mod.WeatherGraphDataset(f, e, window_size=cfg["T"], forecast_horizon=cfg["H"])
        tasks.append((Subset(ds, list(range(n_support))), Subset(ds, list(range(n_support, len(ds)))), {}))
    return tasks


def test_inner_loop_v4_dropin_signature_and_values():
    """inner_loop_v4(model, koppen, support_ds, device): 6 epochs x first 15 windows (here 4 windows)."""
    from weatherforecast_stgcn_maml_b200 import dataset as D
    from weatherforecast_stgcn_maml_b200 import train_hybrid_maml_v5 as TR
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding

    z, cfg, sd, feats, ei, dims = _small_cfg()
    hyb, kop = _hybrid(cfg, sd), KoppenEmbedding(8).cuda()
    support, query, _ = _ref_tasks(cfg, [feats], [ei], 4, D)[0]
    adapted, akop = TR.inner_loop_v4(hyb, kop, support, "cuda")
    assert adapted is not hyb and adapted.training and akop is not kop
    steps = [0, 1, 2, 3] * TR.INNER_EPOCHS_PER_TASK
    fast, _ = P.inner_loop(sd, feats, ei, steps, window=cfg["T"], horizon=cfg["H"], lr=TR.INNER_LR,
                           lstm_layers=cfg["layers"])
    asd = adapted.state_dict()
    for k in sd:
        if k.startswith("base_stgcn."):
            assert torch.equal(asd[k].cpu(), sd[k])
        else:
            assert rel_err(asd[k], fast[k]) <= 2e-4, k
    assert all(torch.equal(a, b) for a, b in zip(hyb.state_dict().values(), [v.cuda() for v in sd.values()]))


def test_meta_update_v4_literal_reference_is_a_noop_and_fomaml_steps():
    from weatherforecast_stgcn_maml_b200 import dataset as D
    from weatherforecast_stgcn_maml_b200 import train_hybrid_maml_v5 as TR
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding

    z, cfg, sd, feats, ei, dims = _small_cfg()
    feats2 = synth.synth_features(feats.shape[0], feats.shape[1], 78)
    tasks = _ref_tasks(cfg, [feats, feats2, feats], [ei, ei, ei], 3, D)
    sched = lambda idx: list(idx[:15])  # one epoch keeps the oracle cheap
    kw = dict(window=cfg["T"], horizon=cfg["H"], lr=TR.INNER_LR, lstm_layers=cfg["layers"])
    # literal reading (D5): loss is reported, no parameter moves, optimiser state stays empty
    hyb, kop = _hybrid(cfg, sd), KoppenEmbedding(8).cuda()
    opt = torch.optim.AdamW(list(hyb.parameters()) + list(kop.parameters()), lr=TR.OUTER_LR, weight_decay=1e-4)
    loss = TR.meta_update_v4(hyb, kop, tasks, "cuda", opt, literal_reference=True, support_schedule=sched)
    ref = sum(float(P.fomaml_task(sd, f, ei, [0, 1, 2], 3, 2, **kw)[0]) for f in (feats, feats2, feats))
    assert abs(loss - ref) <= FWD_TOL * ref
    assert len(opt.state) == 0
    assert all(torch.equal(v.cpu(), sd[k]) for k, v in hyb.state_dict().items())
    # FOMAML reading: groups of 2 tasks, optimiser step after each group (train_hybrid_maml_v5.py:173-179)
    hyb, kop = _hybrid(cfg, sd), KoppenEmbedding(8).cuda()
    opt = torch.optim.AdamW(list(hyb.parameters()) + list(kop.parameters()), lr=TR.OUTER_LR, weight_decay=1e-4)
    loss = TR.meta_update_v4(hyb, kop, tasks, "cuda", opt, support_schedule=sched)
    names = P.trainable(sd)
    cur = {k: v.clone() for k, v in sd.items()}
    params = [torch.nn.Parameter(cur[k].clone()) for k in names]
    ropt = torch.optim.AdamW(params, lr=TR.OUTER_LR, weight_decay=1e-4)
    tot = 0.0
    for group in ([feats, feats2], [feats]):
        gsum = None
        for f in group:
            l, g, _ = P.fomaml_task(cur, f, ei, [0, 1, 2], 3, 2, **kw)
            tot += float(l)
            gsum = g if gsum is None else {k: gsum[k] + g[k] for k in g}
        for p, k in zip(params, names):
            p.grad = gsum[k].clone()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        ropt.step()
        for p, k in zip(params, names):
            cur[k] = p.detach().clone()
    assert abs(loss - tot) <= FWD_TOL * tot
    out = hyb.state_dict()
    for k in names:
        assert rel_err(out[k], cur[k]) <= 1e-4, k
    assert all(p.grad is None for n, p in hyb.named_parameters() if n.startswith("base_stgcn."))
    assert all(p.grad is None for p in kop.parameters())


def test_fine_tune_steps_and_validation_vs_oracle():
    """adapt_hybrid_v5.py:185-231 restated around synthetic data: Adam(L2), clip 1.0, batch-1 steps."""
    from weatherforecast_stgcn_maml_b200.adapt_hybrid_v5 import FineTuner

    z, cfg, sd, feats, ei, dims = _small_cfg()
    feats = synth.synth_features(cfg["T"] + cfg["H"] + 12, feats.shape[1], 5)
    n_win = P.num_windows(feats, cfg["T"], cfg["H"])
    names = P.trainable(sd)
    order = [3, 0, 5, 1, 7, 2]
    for use_graph in (False, True):
        ft = FineTuner(sd, feats, ei, dims, "cuda", dropout=(0, 0, 0), region_name="Thailand", max_samples=n_win, train_frac=0.8,
                       use_cuda_graph=use_graph, val_batch=2)
        assert ft.train_size == int(0.8 * n_win) and abs(ft.initial_lr - 0.0006 * 0.9) < 1e-12
        avg = ft.train_epoch(order)
        val = ft.validate()
        # oracle
        cur = {k: v.clone() for k, v in sd.items()}
        params = [torch.nn.Parameter(cur[k].clone()) for k in names]
        opt = torch.optim.Adam(params, lr=0.0006 * 0.9, weight_decay=1e-5)
        losses = []
        for i in order:
            x, y = P.window_xy(feats, i, cfg["T"], cfg["H"])
            l, g, _ = P.loss_and_grads(cur, x, y, ei, cfg["T"], cfg["H"], 1.0, cfg["layers"])
            for p, k in zip(params, names):
                p.grad = g[k].clone()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            for p, k in zip(params, names):
                cur[k] = p.detach().clone()
            losses.append(float(l))
        assert abs(avg - np.mean(losses)) <= FWD_TOL * np.mean(losses)
        out = ft.state_dict()
        for k in names:
            assert rel_err(out[k], cur[k]) <= 2e-4, (use_graph, k)
        vref = []
        for i in ft.val_idx:
            x, y = P.window_xy(feats, i, cfg["T"], cfg["H"])
            vref.append(float(torch.nn.functional.mse_loss(P.hybrid_forward(cur, x, ei, cfg["T"], cfg["H"], 12, cfg["layers"]), y)))
        assert abs(val - np.mean(vref)) <= 2e-4 * np.mean(vref)


def test_adapt_region_checkpoint_contract():
    from weatherforecast_stgcn_maml_b200.adapt_hybrid_v5 import adapt_region
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding

    z, cfg, sd, feats, ei, dims = _small_cfg()
    feats = synth.synth_features(cfg["T"] + cfg["H"] + 10, feats.shape[1], 6)
    ckpt = {"hybrid_model_state_dict": sd, "koppen_embed_state_dict": KoppenEmbedding(8).state_dict(),
            "config": {"input_channels": cfg["cin"], "hidden_channels": cfg["hidden"], "output_channels": cfg["out"],
                       "window_size": cfg["T"], "forecast_horizon": cfg["H"]},
            "hybrid_config": {"lstm_hidden_size": cfg["L"], "lstm_num_layers": cfg["layers"], "lstm_dropout": 0.2}}
    out = adapt_region(ckpt, feats, ei, (18, 23, 75, 80), "India", stats={"mean": np.zeros(12), "std": np.ones(12)},
                       epochs=2, verbose=False)
    assert set(out) == {"hybrid_model_state_dict", "koppen_embed_state_dict", "region", "region_name", "climate_type",
                        "stats", "config", "hybrid_config", "model_version", "adaptation_type", "val_loss",
                        "base_model_loss", "total_params"}
    assert list(out["hybrid_model_state_dict"].keys()) == list(sd.keys())
    assert out["total_params"] == sum(v.numel() for v in sd.values()) and np.isfinite(out["val_loss"])
    _hybrid(cfg, out["hybrid_model_state_dict"])  # loads back into the drop-in classes


def test_meta_checkpoint_layout_and_real_resume():
    """the meta-training checkpoint carries the reference's keys (train_hybrid_maml_v5.py:311-335),
    its optimiser / scheduler entries load into torch.optim.AdamW / CosineAnnealingWarmRestarts built the reference's way,
    and a trainer resumed from it continues bit-identically (weights, AdamW moments, step count, learning rate)."""
    from weatherforecast_stgcn_maml_b200 import checkpoint as ck
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding
    from weatherforecast_stgcn_maml_b200.schedule import CosineWarmRestarts
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer

    z, cfg, sd, feats, ei, dims = _small_cfg()
    feats2 = synth.synth_features(feats.shape[0], feats.shape[1], 78)
    tasks = [(feats, ei), (feats2, ei)]
    kop = KoppenEmbedding(8)
    mk = lambda state: MetaTrainer(state, tasks, dims, "cuda", dropout=(0, 0, 0), support_rows=(0, 1, 2), query_row=3, accum=2)
    a, sched_a = mk(sd), CosineWarmRestarts(1e-3, 10, 2, 1e-6)
    for _ in range(3):  # three "epochs" of one meta-step each, the schedule stepped per epoch (:294)
        a.meta_step()
        a.set_lr(sched_a.step())
    ckpt = ck.meta_checkpoint(a, kop.state_dict(), sched_a, epoch=3, best_loss=0.5)
    assert set(ckpt) == {"hybrid_model_state_dict", "koppen_embed_state_dict", "meta_optimizer_state_dict", "scheduler_state_dict",
                         "epoch", "best_loss", "model_version", "total_params", "config", "hybrid_config"}
    assert ckpt["model_version"] == "5.0" and ckpt["total_params"] == sum(v.numel() for v in sd.values())
    assert ckpt["config"] == {"input_channels": cfg["cin"], "hidden_channels": cfg["hidden"], "output_channels": cfg["out"],
                              "window_size": cfg["T"], "forecast_horizon": cfg["H"]}
    assert list(ckpt["hybrid_model_state_dict"]) == list(sd)
    # survives torch.save / torch.load the way the reference reads it (adapt_hybrid_v5.py:84)
    import io
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    buf.seek(0)
    ckpt = torch.load(buf, weights_only=False)
    # loads into the reference's optimiser and scheduler objects
    module = _hybrid(cfg, ckpt["hybrid_model_state_dict"])  # same parameter order as the reference's class
    assert [k for k, _ in module.named_parameters()] == list(sd)
    opt = torch.optim.AdamW(list(module.parameters()) + list(kop.parameters()), lr=1e-3, weight_decay=1e-4)
    opt.load_state_dict(ckpt["meta_optimizer_state_dict"])
    ref_sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=10, T_mult=2, eta_min=1e-6)
    ref_sched.load_state_dict({**ref_sched.state_dict(), **ckpt["scheduler_state_dict"]})
    assert abs(ref_sched.get_last_lr()[0] - sched_a.lr) <= 1e-15
    n_state = sum(1 for p in module.parameters() if p in opt.state)
    assert n_state == len(P.trainable(sd))
    # real resume: a fresh trainer from the ORIGINAL weights, then the checkpoint; both continue for two more epochs
    b, sched_b = mk(sd), CosineWarmRestarts(1e-3, 10, 2, 1e-6)
    epoch, best = ck.resume(b, sched_b, ckpt)
    assert (epoch, best) == (3, 0.5) and b.adam.step_count == 3 and b.adam.lr == a.adam.lr
    for _ in range(2):
        la, lb = a.meta_step().item(), b.meta_step().item()
        assert la == lb
        a.set_lr(sched_a.step())
        b.set_lr(sched_b.step())
    torch.cuda.synchronize()
    assert torch.equal(a.theta, b.theta) and torch.equal(a.adam.exp_avg, b.adam.exp_avg)
    assert torch.equal(a.adam.exp_avg_sq, b.adam.exp_avg_sq)
    with pytest.raises(ValueError):
        ck.resume(b, sched_b, {**ckpt, "config": {**ckpt["config"], "window_size": 7}})


def test_task_slots_and_epoch_driver(tmp_path):
    """The reference's main() loop (train_hybrid_maml_v5.py:242-372): BATCH_SIZE tasks sampled per epoch out of all
    regions, one meta-update, cosine schedule, CSV log, best / final checkpoints.  ``MetaTrainer(slots=...)`` keeps all
    tasks resident and re-points its slots between (graph-replayed) meta-steps."""
    from weatherforecast_stgcn_maml_b200 import checkpoint as ck
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding
    from weatherforecast_stgcn_maml_b200.schedule import AdaptiveTaskSampler, CosineWarmRestarts
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer, train_meta

    z, cfg, sd, feats, ei, dims = _small_cfg()
    lats, lons = synth.region_grid(cfg["nlat"], cfg["nlon"])
    eis = [ei, P.knn_edges_ckdtree(lats * 1.7, lons, 4), P.knn_edges_ckdtree(lats, lons * 2.3, 4), ei, ei]
    tasks = [(synth.synth_features(feats.shape[0], feats.shape[1], 100 + i), eis[i]) for i in range(5)]
    kw = dict(dropout=(0, 0, 0), support_rows=(0, 1, 2), query_row=3)
    # (1) a slot trainer pointed at tasks (3, 1) is the plain trainer over those two tasks, bit for bit -- also after the
    # graph has been captured for another assignment
    slot = MetaTrainer(sd, tasks, dims, "cuda", slots=2, **kw)
    assert slot.G == 2 and slot.num_tasks == 5 and slot.active == [0, 1]
    slot.meta_step()
    slot.read_loss()
    slot.load_state_dict(sd)            # back to the initial weights (re-captures), fresh AdamW state
    slot.adam.exp_avg.zero_(); slot.adam.exp_avg_sq.zero_(); slot.adam.step_count = 0
    slot.assign([3, 1])
    plain = MetaTrainer(sd, [tasks[3], tasks[1]], dims, "cuda", **kw)
    for _ in range(2):
        slot.meta_step(); plain.meta_step()
    assert slot.read_loss() == plain.read_loss()
    assert torch.equal(slot.theta, plain.theta)
    with pytest.raises(ValueError):
        slot.assign([0, 1, 2])
    # (2) the epoch driver
    np.random.seed(42)
    tr = MetaTrainer(sd, tasks, dims, "cuda", slots=2, **kw)
    out = train_meta(tr, KoppenEmbedding(8).state_dict(), num_epochs=4, batch_size=2, save_dir=str(tmp_path / "SavedModels"),
                     log_file=str(tmp_path / "log.csv"), verbose=False)
    np.random.seed(42)                   # the draws and the schedule, replayed on their own
    sampler, sched, twin = AdaptiveTaskSampler(5, 2), CosineWarmRestarts(1e-3), MetaTrainer(sd, tasks, dims, "cuda", slots=2, **kw)
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
    ref_sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(ref_opt, T_0=10, T_mult=2, eta_min=1e-6)
    for epoch in range(4):
        ids = sampler.sample()
        twin.assign(ids)
        twin.meta_step()
        loss = twin.read_loss()
        sampler.update(loss)
        sched.step(); ref_opt.step(); ref_sched.step()
        twin.set_lr(sched.get_last_lr()[0])
        assert loss == out["losses"][epoch]
        assert abs(out["lrs"][epoch] - ref_sched.get_last_lr()[0]) <= 1e-15
    assert torch.equal(twin.theta, tr.theta)
    lines = open(tmp_path / "log.csv").read().strip().splitlines()
    assert lines[0] == "epoch,meta_loss,learning_rate" and len(lines) == 5
    assert [float(l.split(",")[1]) for l in lines[1:]] == out["losses"]
    best = torch.load(out["best_path"], weights_only=False)
    final = torch.load(out["final_path"], weights_only=False)
    assert best["best_loss"] == min(out["losses"]) == out["best_loss"] and best["epoch"] == int(np.argmin(out["losses"]))
    assert final["epoch"] == 4 and final["final_loss"] == out["losses"][-1] and "final_loss" not in best
    fresh, fsched = MetaTrainer(sd, tasks, dims, "cuda", slots=2, **kw), CosineWarmRestarts(1e-3)
    epoch, best_loss = ck.resume(fresh, fsched, final)
    assert (epoch, best_loss) == (4, out["best_loss"]) and torch.equal(fresh.theta, tr.theta) and fresh.adam.step_count == 4
