"""Synthetic ERA5-shaped inputs and deterministic v5 weights.

No dataset or checkpoint is reachable offline (the reference's checkpoints are
git-LFS stubs, its data lives on a Windows path), so tests and the benchmark
run on tensors of the reference's shapes:

* region grid: ``lats`` descending / ``lons`` ascending at 0.25 degrees like ERA5
  (graphBuilder.py:23-30 consumes exactly these two 1-D arrays);
* features ``f32[time, N, 24]``: channels 0-11 z-scored weather ~ N(0,1)
  (featurePreprocessor.py:147), 12-15 the four hourly time phases
  (embed_utils.py:18-26, order of featurePreprocessor.py:60-65), 16-23 one
  Koppen-embedding row broadcast to every (time, node)
  (featurePreprocessor.py:170-177);
* weights: the 28 tensors of the v5 hybrid ``state_dict``
  drawn from the reference constructors' distributions with a private
  generator, so both sides of a parity test load the very same numbers.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


class GridCoords:
    """Duck-typed stand-in for the xarray dataset ``build_spatial_graph`` reads
    (graphBuilder.py:23-24 touches only ``.latitude.values`` / ``.longitude.values``)."""

    class _Axis:
        def __init__(self, values):
            self.values = np.asarray(values, dtype=np.float64)

    def __init__(self, lats, lons):
        self.latitude = GridCoords._Axis(lats)
        self.longitude = GridCoords._Axis(lons)


def region_grid(nlat=21, nlon=21, lat0=23.0, lon0=75.0, step=0.25):
    lats = lat0 - step * np.arange(nlat, dtype=np.float64)   # ERA5 latitudes descend
    lons = lon0 + step * np.arange(nlon, dtype=np.float64)
    return lats, lons


def koppen_table(seed=42, num_classes=31, dim=8):
    """N(0,1) table like ``nn.Embedding(31, 8)`` (embed_utils.py:30-35)."""
    g = torch.Generator().manual_seed(seed + 977)
    return torch.randn(num_classes, dim, generator=g)


def synth_features(time_steps, num_nodes, seed, koppen_row=None, channels=24):
    g = torch.Generator().manual_seed(int(seed))
    f = torch.empty(time_steps, num_nodes, channels, dtype=torch.float32)
    f[..., :12] = torch.randn(time_steps, num_nodes, 12, generator=g)
    hours = torch.arange(time_steps, dtype=torch.float64)
    doy = 1.0 + torch.floor(hours / 24.0)
    yp = 2.0 * math.pi * doy / 365.25
    dp = 2.0 * math.pi * (hours % 24.0) / 24.0
    tfeat = torch.stack([yp.sin(), yp.cos(), dp.sin(), dp.cos()], dim=-1).to(torch.float32)
    f[..., 12:16] = tfeat[:, None, :]
    if koppen_row is None:
        koppen_row = torch.zeros(channels - 16)
    f[..., 16:] = koppen_row.to(torch.float32)[None, None, :]
    return f.contiguous()


def synth_task(task_id, num_windows=600, nlat=21, nlon=21, window=24, horizon=8, base_seed=42):
    """One synthetic region task: (lats, lons, features, koppen_code)."""
    lats, lons = region_grid(nlat, nlon, lat0=23.0 + 5.0 * (task_id % 7), lon0=75.0 + 5.0 * (task_id % 11))
    code = 1 + task_id % 30
    row = koppen_table(base_seed)[code]
    feats = synth_features(num_windows + window + horizon, nlat * nlon, base_seed + task_id, row)
    return lats, lons, feats, code


def v5_shapes(in_channels=24, hidden=256, lstm_hidden=128, lstm_layers=4, out_channels=12, horizon=8):
    """(name, shape) of the hybrid ``state_dict`` in registration order."""
    s = []
    for i in range(1, 5):
        cin = in_channels if i == 1 else hidden
        s.append((f"base_stgcn.conv{i}.bias", (hidden,)))
        s.append((f"base_stgcn.conv{i}.lin.weight", (hidden, cin)))
    s.append(("base_stgcn.output_layer.weight", (out_channels * horizon, hidden)))
    s.append(("base_stgcn.output_layer.bias", (out_channels * horizon,)))
    for l in range(lstm_layers):
        kin = hidden if l == 0 else lstm_hidden
        s.append((f"lstm.weight_ih_l{l}", (4 * lstm_hidden, kin)))
        s.append((f"lstm.weight_hh_l{l}", (4 * lstm_hidden, lstm_hidden)))
        s.append((f"lstm.bias_ih_l{l}", (4 * lstm_hidden,)))
        s.append((f"lstm.bias_hh_l{l}", (4 * lstm_hidden,)))
    s.append(("output_layer.weight", (out_channels * horizon, lstm_hidden)))
    s.append(("output_layer.bias", (out_channels * horizon,)))
    return s


def init_v5_state_dict(seed=42, gcn_bias_scale=0.0, **cfg):
    """Deterministic weights with the reference constructors' distributions.

    GCN ``lin.weight``: glorot-uniform, ``bias``: zeros (PyG GCNConv); LSTM:
    U(+-1/sqrt(hidden)) (torch nn.LSTM, hybrid_model.py:42-49); Linear:
    U(+-1/sqrt(fan_in)) (model.py:28, hybrid_model.py:52-55).
    ``gcn_bias_scale`` > 0 draws non-zero GCN biases so tests exercise the bias path.
    """
    g = torch.Generator().manual_seed(int(seed))
    lstm_hidden = cfg.get("lstm_hidden", 128)
    sd = OrderedDict()
    for name, shape in v5_shapes(**cfg):
        t = torch.empty(*shape, dtype=torch.float32)
        if name.startswith("base_stgcn.conv") and name.endswith("lin.weight"):
            a = math.sqrt(6.0 / (shape[0] + shape[1]))
        elif name.startswith("base_stgcn.conv"):
            a = gcn_bias_scale
        elif name.startswith("lstm."):
            a = 1.0 / math.sqrt(lstm_hidden)
        elif name.endswith("weight"):
            a = 1.0 / math.sqrt(shape[1])
        else:  # Linear bias: fan_in of the matching weight
            a = 1.0 / math.sqrt(sd[name[:-4] + "weight"].shape[1])
        if a == 0.0:
            t.zero_()
        else:
            t.uniform_(-a, a, generator=g)
        sd[name] = t
    return sd


TRAINABLE_PREFIXES = ("lstm.", "output_layer.")


def trainable_names(sd):
    """The 18 tensors autograd reaches on the hybrid path."""
    return [k for k in sd if k.startswith(TRAINABLE_PREFIXES)]
