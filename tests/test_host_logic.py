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
