"""Drop-in for the reference's ``dataset.WeatherGraphDataset`` (dataset.py:6-54) plus the two
PyG containers the drivers touch (``Data``, ``DataLoader``; torch_geometric is not required).

Window layout contract: sample ``i`` has
``x = features[i : i+W].reshape(W*N, C)`` -- a contiguous slice, rows time-major -- and
``y = features[i+W+1 : i+W+1+H, :, :12].reshape(H*N, 12)`` (index ``i+W`` itself is skipped).
``window_offsets`` exposes the same windows as element offsets into the resident features
tensor, which is how the batched engine reads them with zero copies.
"""
from __future__ import annotations

import torch
import torch.utils.data
from torch.utils.data import Dataset


class Data:
    """Minimal ``torch_geometric.data.Data``: attribute bag with ``.to(device)``."""

    def __init__(self, x=None, edge_index=None, y=None):
        self.x, self.edge_index, self.y = x, edge_index, y

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device, non_blocking=False):
        return Data(self.x.to(device, non_blocking=non_blocking), self.edge_index.to(device, non_blocking=non_blocking),
                    self.y.to(device, non_blocking=non_blocking))


def _collate(items):
    if len(items) == 1:
        return items[0]
    xs, eis, ys, off = [], [], [], 0
    for d in items:  # PyG batching: concatenate nodes, offset edge ids
        xs.append(d.x)
        ys.append(d.y)
        eis.append(d.edge_index + off)
        off += d.num_nodes
    return Data(torch.cat(xs), torch.cat(eis, dim=1), torch.cat(ys))


class DataLoader(torch.utils.data.DataLoader):
    """``torch_geometric.loader.DataLoader`` for ``Data`` items."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, collate_fn=_collate, **kwargs)


class WeatherGraphDataset(Dataset):
    def __init__(self, features, edge_index, window_size=6, forecast_horizon=1):
        self.features = features
        self.edge_index = edge_index
        self.window_size = window_size
        self.forecast_horizon = forecast_horizon
        self.num_weather_vars = 12
        self.num_nodes = features.shape[1]
        self.valid_indices = range(window_size, len(features) - forecast_horizon)

    def __len__(self):
        return len(self.valid_indices)

    def __getitem__(self, idx):
        actual_idx = self.valid_indices[idx]
        start = actual_idx - self.window_size
        x = self.features[start:actual_idx].reshape(self.window_size * self.num_nodes, -1)
        first = actual_idx + 1
        y = self.features[first:first + self.forecast_horizon, :, : self.num_weather_vars]
        y = y.reshape(self.forecast_horizon * self.num_nodes, self.num_weather_vars)
        return Data(x=x.clone().detach(), edge_index=self.edge_index, y=y.clone().detach())

    # ---- zero-copy view of the same windows ------------------------------------------------
    def window_offsets(self, idx):
        """(x_offset, target_offset) in elements into ``features`` for sample ``idx``."""
        per_step = self.num_nodes * self.features.shape[2]
        start = self.valid_indices[idx] - self.window_size
        return start * per_step, (start + self.window_size + 1) * per_step

    def time_span(self, idx):
        """[first, last) time rows sample ``idx`` reads (inputs and targets)."""
        start = self.valid_indices[idx] - self.window_size
        return start, start + self.window_size + 1 + self.forecast_horizon


def unwrap_subset(ds):
    """(WeatherGraphDataset, [indices]) for a dataset or (nested) ``torch.utils.data.Subset``."""
    idx = None
    while isinstance(ds, torch.utils.data.Subset):
        idx = list(ds.indices) if idx is None else [ds.indices[i] for i in idx]
        ds = ds.dataset
    if idx is None:
        idx = list(range(len(ds)))
    return ds, idx
