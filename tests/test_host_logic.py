"""CPU tests of the host-side logic: dataset layout, schedules, staging plan, schedulers, sharding."""
import math

import numpy as np
import pytest
import torch

from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth


def test_dataset_matches_oracle_windows_and_offsets():
    from weatherforecast_stgcn_maml_b200.dataset import DataLoader, WeatherGraphDataset, unwrap_subset
    from torch.utils.data import Subset

    feats = synth.synth_features(60, 6, 0)
    ei = torch.zeros(2, 0, dtype=torch.long)
    ds = WeatherGraphDataset(feats, ei, window_size=24, forecast_horizon=8)
    assert len(ds) == 28 == P.num_windows(feats, 24, 8)
    flat = feats.reshape(-1)
    for i in (0, 5, 27):
        d = ds[i]
        x, y = P.window_xy(feats, i, 24, 8)
        assert torch.equal(d.x, x) and torch.equal(d.y, y)
        xo, to = ds.window_offsets(i)
        assert torch.equal(flat[xo:xo + x.numel()].view_as(x), x)
        assert torch.equal(flat[to:to + 8 * 6 * 24].view(8 * 6, 24)[:, :12], y)
        assert ds.time_span(i) == (i, i + 33)
    inner, idx = unwrap_subset(Subset(Subset(ds, [4, 5, 6, 7]), [1, 3]))
    assert inner is ds and idx == [5, 7]
    batch = next(iter(DataLoader(Subset(ds, [2]), batch_size=1, shuffle=False)))
    assert torch.equal(batch.x, ds[2].x)
    two = next(iter(DataLoader(ds, batch_size=2)))
    assert two.x.shape[0] == 2 * 24 * 6


def test_reference_support_schedule():
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import (INNER_EPOCHS_PER_TASK,
                                                                        reference_support_schedule)

    s = reference_support_schedule(list(range(450)))
    assert len(s) == 90 == INNER_EPOCHS_PER_TASK * 15 and s[:15] == list(range(15)) and s[15:30] == list(range(15))
    assert reference_support_schedule([7, 8, 9]) == [7, 8, 9] * 6


def test_constants_match_reference():
    from weatherforecast_stgcn_maml_b200 import train_hybrid_maml_v5 as TR

    assert (TR.INNER_LR, TR.OUTER_LR, TR.GRAD_ACCUMULATION_STEPS, TR.WINDOW_SIZE, TR.FORECAST_HORIZON) == (0.01, 0.001, 2, 24, 8)
    assert (TR.HIDDEN_CHANNELS, TR.LSTM_HIDDEN_SIZE, TR.LSTM_NUM_LAYERS, TR.INPUT_CHANNELS) == (256, 128, 4, 24)


def test_climate_scheduler_matches_reference_formula():
    from weatherforecast_stgcn_maml_b200.adaptive_scheduler import (ClimateAwareLRScheduler, climate_hyperparameters,
                                                                     create_climate_optimizer)

    assert climate_hyperparameters("Thailand") == (0.0006 * 0.9, 1e-5)
    assert climate_hyperparameters("Moscow") == (0.0006 * 1.1, 5e-5)
    assert climate_hyperparameters("Paris") == (0.0006, 1e-4)
    p = torch.nn.Parameter(torch.zeros(3))
    opt, lr = create_climate_optimizer([p], "Moscow")
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["weight_decay"] == 5e-5
    sch = ClimateAwareLRScheduler(opt, "Moscow", lr)
    got = [sch.step(l) for l in (2.0, 2.0, 2.0, 2.0, 0.1, 0.5)]
    exp = []
    for ep, l in enumerate((2.0, 2.0, 2.0, 2.0, 0.1, 0.5), 1):
        v = lr * 1.1 * 0.5 * (1 + math.cos(math.pi * ((ep - 1) % 5 / 5)))
        if ep > 3:
            v *= 1.1 if l > 1.0 else (0.95 if l < 0.2 else 1.0)
        exp.append(v)
    assert np.allclose(got, exp, rtol=0, atol=1e-15) and sch.get_last_lr() == [exp[-1]]


def test_trainable_layout_matches_state_dict_order():
    from weatherforecast_stgcn_maml_b200.engine import V5Dims, flatten_trainable, trainable_layout, unflatten_trainable

    dims = V5Dims()
    lay = trainable_layout(dims)
    sd = synth.init_v5_state_dict(1)
    assert [n for n, _, _ in lay] == synth.trainable_names(sd)
    assert lay[-1][2] + 96 == 606304 == dims.P
    flat = flatten_trainable(sd, dims)
    back = unflatten_trainable(flat, dims)
    assert all(torch.equal(back[k], sd[k]) for k in back)


def test_task_sharding_round_robin():
    from weatherforecast_stgcn_maml_b200.dist import shard_tasks

    assert shard_tasks(15, 0, 1) == list(range(15))
    parts = [shard_tasks(120, r, 8) for r in range(8)]
    assert all(len(p) == 15 for p in parts) and sorted(sum(parts, [])) == list(range(120))
    assert shard_tasks(5, 1, 2) == [1, 3]
    with pytest.raises(ValueError):
        shard_tasks(3, 0, 4, require_even=True)


def test_cosine_warm_restarts_matches_torch():
    from weatherforecast_stgcn_maml_b200.schedule import CosineWarmRestarts

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=10, T_mult=2, eta_min=1e-6)
    mine = CosineWarmRestarts(1e-3, T_0=10, T_mult=2, eta_min=1e-6)
    for _ in range(75):  # 40 epochs in the reference (train_hybrid_maml_v5.py:24); run through three restarts
        opt.step()
        ref.step()
        assert mine.step() == pytest.approx(ref.get_last_lr()[0], rel=1e-12, abs=1e-18)


def test_adaptive_task_sampler_follows_the_reference_draws():
    from weatherforecast_stgcn_maml_b200.schedule import AdaptiveTaskSampler

    np.random.seed(42)
    s = AdaptiveTaskSampler(15, 4)
    got = []
    for loss in (0.9, 0.7, 0.65):
        got.append(s.sample())
        s.update(loss)
    # the same sequence written out the way train_hybrid_maml_v5.py:264-292 does it
    np.random.seed(42)
    losses, want = [], []
    for loss in (0.9, 0.7, 0.65):
        if losses:
            want.append(list(np.random.choice(15, 4, replace=False, p=np.array(losses) / sum(losses))))
        else:
            want.append(list(np.random.choice(15, 4, replace=False)))
        losses = [loss] * 15 if len(losses) < 15 else [0.9 * t + 0.1 * loss for t in losses]
    assert got == [[int(i) for i in w] for w in want]
    assert len(set(s.task_losses)) == 1  # every task carries the same "difficulty": the draw is uniform
    assert AdaptiveTaskSampler(3, 4).sample() == [0, 1, 2]


def test_denormalisation_and_forecast_metrics():
    from weatherforecast_stgcn_maml_b200.featurePreprocessor import (denormalize_all_predictions, denormalize_predictions,
                                                                      forecast_metrics)

    rng = np.random.default_rng(0)
    stats = {"mean": rng.normal(size=12) * 10, "std": rng.uniform(0.5, 3.0, size=12)}
    pred = rng.normal(size=(40, 12))
    assert np.allclose(denormalize_all_predictions(pred, stats), pred * stats["std"] + stats["mean"])
    assert np.allclose(denormalize_all_predictions(pred[0], stats), pred[0] * stats["std"] + stats["mean"])
    t = torch.tensor(pred[:, 2], dtype=torch.float32)
    assert torch.allclose(denormalize_predictions(t, stats), t * float(stats["std"][2]) + float(stats["mean"][2]))
    assert denormalize_predictions(t, {}) is t
    H, N = 8, 5
    yp, yt = rng.normal(size=(H * N, 12)), rng.normal(size=(H * N, 12))
    m = forecast_metrics(torch.tensor(yp), yt, stats, N, H)
    a = (yp.reshape(H, N, 12).mean(1) * stats["std"] + stats["mean"])
    b = (yt.reshape(H, N, 12).mean(1) * stats["std"] + stats["mean"])
    assert m["t2m"]["mse"] == pytest.approx(np.mean((a[:, 2] - b[:, 2]) ** 2))
    assert m["sp"]["mae"] == pytest.approx(np.mean(np.abs(a[:, 4] - b[:, 4])))   # sp is channel 4 (featurePreprocessor.py:42-55)
    assert m["average_mse"] == pytest.approx(np.mean([np.mean((a[:, v] - b[:, v]) ** 2) for v in (0, 1, 2, 3, 5)]))


def test_scheduler_state_roundtrips_with_torch():
    """checkpoint.scheduler_state_dict / load_scheduler_state_dict against torch's CosineAnnealingWarmRestarts (the
    object the reference saves, train_hybrid_maml_v5.py:250-252,318): a state written by either side resumes on the
    other and the learning rates stay equal."""
    from weatherforecast_stgcn_maml_b200 import checkpoint as ck
    from weatherforecast_stgcn_maml_b200.schedule import CosineWarmRestarts

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=10, T_mult=2, eta_min=1e-6)
    mine = CosineWarmRestarts(1e-3, 10, 2, 1e-6)
    for _ in range(13):
        opt.step(); ref.step(); mine.step()
    # torch -> here
    resumed = CosineWarmRestarts(5.0, 3, 1, 0.0)
    ck.load_scheduler_state_dict(resumed, ref.state_dict())
    # here -> torch
    opt2 = torch.optim.AdamW([p], lr=1e-3)
    ref2 = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt2, T_0=10, T_mult=2, eta_min=1e-6)
    ref2.load_state_dict({**ref2.state_dict(), **ck.scheduler_state_dict(mine)})
    for _ in range(25):
        opt.step(); opt2.step(); ref.step(); ref2.step()
        a, b = mine.step(), resumed.step()
        assert abs(a - ref.get_last_lr()[0]) <= 1e-12 and abs(b - a) <= 1e-15 and abs(ref2.get_last_lr()[0] - a) <= 1e-12


def test_optimizer_state_dict_loads_into_torch_adamw():
    """The fused optimiser's flat moments, written in torch.optim.AdamW's state_dict format over the reference's parameter
    list (hybrid parameters then the Koppen table, train_hybrid_maml_v5.py:245-249), load into a real AdamW; entries land
    on the right parameters; frozen parameters have no state; the inverse conversion returns the flat buffers; one more
    torch step from the loaded state equals the AdamW recurrence applied to the flat buffers."""
    from types import SimpleNamespace

    from weatherforecast_stgcn_maml_b200 import checkpoint as ck, synth
    from weatherforecast_stgcn_maml_b200.engine import V5Dims, flatten_trainable, trainable_layout

    dims = V5Dims(num_nodes=12, window=6, horizon=2, hidden=32, lstm_hidden=32, lstm_layers=2)
    sd = synth.init_v5_state_dict(7, hidden=32, lstm_hidden=32, lstm_layers=2, horizon=2)
    layout = trainable_layout(dims)
    P = layout[-1][2] + int(torch.Size(layout[-1][1]).numel())
    g = torch.Generator().manual_seed(1)
    adam = SimpleNamespace(exp_avg=torch.randn(P, generator=g), exp_avg_sq=torch.rand(P, generator=g), step_count=5, lr=7e-4,
                           betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, decoupled=True)
    osd = ck.optimizer_state_dict(adam, sd.keys(), dims, num_extra_params=1)
    names = [n for n, _, _ in layout]
    keys = list(sd.keys())
    assert sorted(osd["state"]) == sorted(keys.index(n) for n in names)           # only trainable parameters carry state
    assert osd["param_groups"][0]["params"] == list(range(len(keys) + 1))          # + embedding.weight
    params = [torch.nn.Parameter(v.clone().float()) for v in sd.values()] + [torch.nn.Parameter(torch.zeros(31, 8))]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-4)
    opt.load_state_dict(osd)
    assert opt.param_groups[0]["lr"] == 7e-4
    for n, shape, off in layout:
        st = opt.state[params[keys.index(n)]]
        assert float(st["step"]) == 5 and torch.equal(st["exp_avg"].reshape(-1), adam.exp_avg[off:off + st["exp_avg"].numel()])
    # one torch step from the loaded state == the AdamW recurrence on the flat buffers
    grads = torch.randn(P, generator=g) * 1e-2
    theta = flatten_trainable(sd, dims).clone()
    for n, shape, off in layout:
        params[keys.index(n)].grad = grads[off:off + int(torch.Size(shape).numel())].view(shape).clone()
    opt.step()
    b1, b2, t = 0.9, 0.999, 6
    m = b1 * adam.exp_avg + (1 - b1) * grads
    v = b2 * adam.exp_avg_sq + (1 - b2) * grads * grads
    want = theta * (1 - 7e-4 * 1e-4) - 7e-4 * (m / (1 - b1 ** t)) / ((v / (1 - b2 ** t)).sqrt() + 1e-8)
    got = torch.cat([params[keys.index(n)].detach().reshape(-1) for n in names])
    assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max())
    # inverse conversion (what a resume does), also from torch's own state_dict
    back = SimpleNamespace(exp_avg=torch.zeros(P), exp_avg_sq=torch.zeros(P), step_count=0, lr=0.0, betas=None, eps=0.0,
                           weight_decay=0.0, decoupled=True)
    ck.load_optimizer_state_dict(back, opt.state_dict(), sd.keys(), dims)
    assert back.step_count == 6 and back.lr == 7e-4 and tuple(back.betas) == (0.9, 0.999)
    assert float((back.exp_avg - m).abs().max()) <= 1e-6 and float((back.exp_avg_sq - v).abs().max()) <= 1e-6  # torch lerps


def test_feature_assembly_has_no_cpu_fallback():
    """The product path fails loudly without a CUDA tensor (no silent numpy / torch fallback)."""
    from weatherforecast_stgcn_maml_b200.featurePreprocessor import assemble_features, feature_stats

    w = torch.zeros(4, 3, 12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        feature_stats(w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        assemble_features(w, np.zeros((4, 4), dtype=np.float32), np.zeros(8, dtype=np.float32))


def test_accumulation_groups_follow_the_reference_loop():
    """train_hybrid_maml_v5.py:151-179: which tasks contribute to which optimiser step, including the reference's
    behaviour for missing (None) tasks at boundary positions."""
    import itertools

    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import reference_accumulation_groups

    def reference_loop(tasks, accum):
        steps, pending = [], []
        for i, (support, _q, _s) in enumerate(tasks):
            if support is None:
                continue
            pending.append(support)
            if (i + 1) % accum == 0 or i == len(tasks) - 1:
                steps.append((pending, True))
                pending = []
        if pending:
            steps.append((pending, False))  # gradients computed, never applied
        return steps

    for n in range(0, 7):
        for accum in (1, 2, 3):
            for present in itertools.product([True, False], repeat=n):
                tasks = [(f"s{i}" if p else None, f"q{i}", {}) for i, p in enumerate(present)]
                got = [([t[0] for t in g], step) for g, step in reference_accumulation_groups(tasks, accum)]
                assert got == reference_loop(tasks, accum), (present, accum)
    # the all-present case is plain chunks of `accum` with a step after each
    tasks = [(f"s{i}", None, {}) for i in range(5)]
    assert [([t[0] for t in g], s) for g, s in reference_accumulation_groups(tasks, 2)] == [
        (["s0", "s1"], True), (["s2", "s3"], True), (["s4"], True)]
