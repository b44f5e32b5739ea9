"""CPU restatement of the reference's v5 hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional style: every routine takes the weights as a plain ``dict`` of
tensors keyed like the reference ``state_dict`` instead of
``nn.Module`` objects.  Each function cites the reference lines it follows
(paths are relative to /root/reference).  ``oracle/make_golden.py`` checks this
file against the unmodified reference in the build container and freezes the
results under tests/golden/; tests/test_oracle_golden.py re-checks it anywhere.

Third-party arithmetic used natively (installed in the image): scipy cKDTree,
torch CPU matmul / autograd / optimizers.  ``torch_geometric.GCNConv`` is absent
and restated in ``gcn_norm`` / ``gcn_conv`` (PyG >= 2.0 defaults).
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- graph
def node_positions(lats, lons):
    """graphBuilder.py:27-30 -- meshgrid(indexing='ij'), node id = ilat * nlon + ilon."""
    la, lo = np.meshgrid(np.asarray(lats, np.float64), np.asarray(lons, np.float64), indexing="ij")
    return np.c_[la.ravel(), lo.ravel()]


def knn_edges_ckdtree(lats, lons, k):
    """graphBuilder.py:34-44 -- cKDTree.query(k+1), drop column 0 (self), edges [node, nbr]."""
    from scipy.spatial import cKDTree

    pos = node_positions(lats, lons)
    _, nb = cKDTree(pos).query(pos, k=k + 1)
    n = len(pos)
    src = np.repeat(np.arange(n, dtype=np.int64), k)
    dst = nb[:, 1:].astype(np.int64).ravel()
    return torch.from_numpy(np.stack([src, dst])).contiguous()


def knn_edges_canonical(lats, lons, k):
    """Brute-force k nearest by (squared distance, index) -- the deterministic tie rule.

    Same metric as graphBuilder.py:34-35 (Euclidean on raw degrees, float64).  Where no
    tie straddles the k-th place this equals cKDTree's neighbour *set*; the distance
    multiset is equal everywhere.
    """
    pos = node_positions(lats, lons)
    n = len(pos)
    out = np.empty((n, k), dtype=np.int64)
    for i in range(n):
        dlat = pos[:, 0] - pos[i, 0]
        dlon = pos[:, 1] - pos[i, 1]
        d2 = dlat * dlat + dlon * dlon
        d2[i] = -1.0
        order = np.lexsort((np.arange(n), d2))
        out[i] = order[1 : k + 1]
    src = np.repeat(np.arange(n, dtype=np.int64), k)
    return torch.from_numpy(np.stack([src, out.ravel()])).contiguous()


def knn_sq_distances(lats, lons, edge_index):
    pos = node_positions(lats, lons)
    d = pos[edge_index[0].numpy()] - pos[edge_index[1].numpy()]
    return (d * d).sum(1)


def gcn_norm(edge_index, num_rows):
    """PyG gcn_norm as invoked by every GCNConv.forward (model.py:31-40, hybrid_model.py:65-74).

    num_rows = x.size(0) = window * N, NOT the N of the graph: rows >= N
    only get their self loop, deg 1, weight 1.
    """
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loops = torch.arange(num_rows, dtype=torch.long)
    row = torch.cat([row[keep], loops])
    col = torch.cat([col[keep], loops])
    deg = torch.zeros(num_rows, dtype=torch.float32).scatter_add_(0, col, torch.ones(row.numel()))
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    w = dis[row] * torch.ones(row.numel()) * dis[col]
    return row, col, w


def gcn_conv(x, edge_index, weight, bias):
    """PyG GCNConv.forward: out[i] = sum_{e: dst_e = i} w_e * (x W^T)[src_e] + b."""
    row, col, w = gcn_norm(edge_index, x.size(0))
    h = x @ weight.t()
    out = torch.zeros_like(h).index_add_(0, col, w[:, None] * h[row])
    return out + bias


def gcn_stack(sd, x, edge_index, prefix="base_stgcn.", masks=None):
    """hybrid_model.py:60-78 / model.py:31-42: 4 x relu(GCNConv), each followed by nn.Dropout when ``masks`` gives
    the layer a keep-mask already scaled by 1/(1-p) (``F.dropout(h) == h * mask``); None / missing entries = no dropout
    (eval mode, or the hybrid path's last layer :76)."""
    h = x
    for i in range(1, 5):
        h = torch.relu(gcn_conv(h, edge_index, sd[f"{prefix}conv{i}.lin.weight"], sd[f"{prefix}conv{i}.bias"]))
        if masks is not None and i - 1 < len(masks) and masks[i - 1] is not None:
            h = h * masks[i - 1]
    return h


# --------------------------------------------------------------------------- LSTM + head
def lstm_last_hidden(sd, seq, num_layers, prefix="lstm.", masks=None):
    """torch.nn.LSTM(batch_first, zero initial state) restated cell by cell.

    hybrid_model.py:42-49 (definition) and :93-105 (use): gates ordered i, f, g, o;
    c = f*c + i*g; h = o*tanh(c); returns the top layer's h at the last step,
    i.e. ``lstm_out[:, -1, :]``.  ``seq`` is [batch, T, F]: the reference feeds one node
    at a time (batch 1); batching over nodes is the same arithmetic per row.  ``masks[l]`` ([batch, T, hid], scaled
    keep-mask): nn.LSTM's inter-layer dropout on the output of layer l < num_layers - 1 (hybrid_model.py:47).
    """
    inp = seq
    for l in range(num_layers):
        w_ih, w_hh = sd[f"{prefix}weight_ih_l{l}"], sd[f"{prefix}weight_hh_l{l}"]
        b = sd[f"{prefix}bias_ih_l{l}"] + sd[f"{prefix}bias_hh_l{l}"]
        hid = w_hh.shape[1]
        h = seq.new_zeros(seq.shape[0], hid)
        c = seq.new_zeros(seq.shape[0], hid)
        outs = []
        for t in range(inp.shape[1]):
            g = inp[:, t] @ w_ih.t() + h @ w_hh.t() + b
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        if masks is not None and l < num_layers - 1 and masks[l] is not None:
            inp = inp * masks[l]
    return inp[:, -1]


def hybrid_forward(sd, x, edge_index, window, horizon=8, out_channels=12, lstm_layers=4, masks=None):
    """hybrid_model.py:80-117.  ``masks`` = None: eval mode (dropout off).  Train mode with KNOWN masks (each already
    scaled by 1/(1-p)): {"gcn": [3 x [T*N, hidden]] (:67-73), "lstm": [layers-1 x [N, T, L]] (:47), "head": [N, L] (:108)}.
    Returns f32[N*horizon, out]."""
    masks = masks or {}
    with torch.no_grad():  # hybrid_model.py:63 -- unconditional, even when not frozen (D4)
        feats = gcn_stack({k: v.detach() for k, v in sd.items() if k.startswith("base_stgcn.")}, x, edge_index,
                          masks=masks.get("gcn"))
    n = feats.shape[0] // window
    seq = feats.view(window, n, -1).permute(1, 0, 2)  # :89-90
    last = lstm_last_hidden(sd, seq, lstm_layers, masks=masks.get("lstm"))  # :93-105
    if masks.get("head") is not None:
        last = last * masks["head"]  # :108
    pred = last @ sd["output_layer.weight"].t() + sd["output_layer.bias"]  # :111
    return pred.view(n, horizon, out_channels).reshape(-1, out_channels)  # :114-115 (row = node*H + h)


def stgcn_forward(sd, x, edge_index, window, horizon=8, out_channels=12, prefix="", masks=None):
    """model.py:30-52: differentiable 4 x GCN (+ dropout after each when ``masks`` lists four scaled keep-masks; None =
    eval mode), last N rows, Linear head."""
    h = gcn_stack(sd, x, edge_index, prefix=prefix, masks=masks)
    n = h.shape[0] // window
    h = h[-n:]
    out = h @ sd[f"{prefix}output_layer.weight"].t() + sd[f"{prefix}output_layer.bias"]
    return out.view(n, horizon, out_channels).reshape(-1, out_channels)


# --------------------------------------------------------------------------- windows
def window_xy(features, idx, window=24, horizon=8, num_weather=12):
    """dataset.py:30-48: x = features[idx : idx+W] flattened time-major; targets at
    actual_idx + h for h = 1..H with actual_idx = idx + W (index actual_idx itself is skipped)."""
    n = features.shape[1]
    x = features[idx : idx + window].reshape(window * n, -1)
    a = idx + window
    y = torch.stack([features[a + h, :, :num_weather] for h in range(1, horizon + 1)], 0)
    return x.clone(), y.reshape(horizon * n, num_weather).clone()


def num_windows(features, window=24, horizon=8):
    """dataset.py:25: len(range(window, T - horizon))."""
    return features.shape[0] - window - horizon


# --------------------------------------------------------------------------- training pieces
def trainable(sd):
    return [k for k in sd if k.startswith(("lstm.", "output_layer."))]


def clip_grad_norm(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ (train_hybrid_maml_v5.py:135-138): one global L2
    norm over the non-None grads, scale by min(1, max_norm / (norm + 1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def loss_and_grads(sd, x, y, edge_index, window, horizon=8, scale=1.0, lstm_layers=4, masks=None):
    """forward + nn.MSELoss + backward (train_hybrid_maml_v5.py:132-134, 166-169).

    Returns (loss, {name: grad}) over the 18 tensors autograd reaches (D4)."""
    names = trainable(sd)
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in names else v.detach()) for k, v in sd.items()}
    pred = hybrid_forward(leaf, x, edge_index, window, horizon, y.shape[1], lstm_layers, masks=masks)
    loss = F.mse_loss(pred, y) * scale
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), dict(zip(names, grads)), pred.detach()


def inner_loop(sd, features, edge_index, steps, window=24, horizon=8, lr=0.01, lstm_layers=4):
    """train_hybrid_maml_v5.py:110-141: deepcopy, then for each listed support window:
    zero_grad, forward, MSE, backward, clip(1.0), plain SGD.  ``steps`` is the sequence of
    support-window indices (reference: 6 x [0..14])."""
    fast = {k: v.clone() for k, v in sd.items()}
    losses = []
    for idx in steps:
        x, y = window_xy(features, idx, window, horizon)
        loss, grads, _ = loss_and_grads(fast, x, y, edge_index, window, horizon, 1.0, lstm_layers)
        gl = [grads[k].clone() for k in grads]
        clip_grad_norm(gl, 1.0)
        for k, g in zip(grads, gl):
            fast[k] = fast[k] - lr * g
        losses.append(float(loss))
    return fast, losses


def reference_support_schedule_indices(num_support, epochs=6, per_epoch=15):
    """Window order of inner_loop_v4 (train_hybrid_maml_v5.py:124-127): ``epochs`` passes over the first ``per_epoch``
    support windows, in order (shuffle=False, ``break`` at batch 15)."""
    return list(range(min(num_support, per_epoch))) * epochs


def fomaml_task(sd, features, edge_index, support_steps, query_idx, accum, **kw):
    """One task of meta_update_v4 (train_hybrid_maml_v5.py:151-170): inner loop, then the
    first query window, loss / GRAD_ACCUMULATION_STEPS, backward INTO THE ADAPTED COPY.
    Returns (query_loss/accum, grads on the adapted copy, adapted weights)."""
    fast, _ = inner_loop(sd, features, edge_index, support_steps, **kw)
    x, y = window_xy(features, query_idx, kw.get("window", 24), kw.get("horizon", 8))
    loss, grads, _ = loss_and_grads(fast, x, y, edge_index, kw.get("window", 24), kw.get("horizon", 8),
                                    1.0 / accum, kw.get("lstm_layers", 4))
    return loss, grads, fast


# --------------------------------------------------------------------------- literal-cost model
def build_reference_like_module(sd, window, horizon=8, lstm_layers=4):
    """An nn.LSTM + per-node Python loop with the reference's execution shape
    (hybrid_model.py:93-105: one nn.LSTM call per node, batch 1).  Used by bench.py's
    cpu_baseline leg so the timed CPU work has the reference's cost profile."""
    hid = sd["lstm.weight_hh_l0"].shape[1]
    lstm = torch.nn.LSTM(sd["lstm.weight_ih_l0"].shape[1], hid, lstm_layers, batch_first=True)
    with torch.no_grad():
        for l in range(lstm_layers):
            for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                getattr(lstm, f"{nm}_l{l}").copy_(sd[f"lstm.{nm}_l{l}"])
    head_w = sd["output_layer.weight"].clone().requires_grad_(True)
    head_b = sd["output_layer.bias"].clone().requires_grad_(True)

    def forward(x, edge_index):
        with torch.no_grad():
            feats = gcn_stack(sd, x, edge_index)
        n = feats.shape[0] // window
        seq = feats.view(window, n, -1).permute(1, 0, 2)
        outs = []
        for node in range(n):
            o, _ = lstm(seq[node : node + 1])
            outs.append(o[0, -1, :])
        last = torch.stack(outs, 0)
        pred = last @ head_w.t() + head_b
        return pred.view(n, horizon, -1).reshape(-1, pred.shape[1] // horizon)

    params = list(lstm.parameters()) + [head_w, head_b]
    return forward, params


# ------------------------------------------------------------------ feature assembly
WEATHER_VARS = ["u10", "v10", "t2m", "d2m", "sp", "tp", "u100", "v100", "str", "hcc", "lcc", "e"]
TIME_VARS = ["year_progress_sin", "year_progress_cos", "day_progress_sin", "day_progress_cos"]


def time_features(day_of_year, time_of_day):
    """embed_utils.py:12-26 on plain arrays: [time, 4] f64 in TIME_VARS order."""
    year = 2 * np.pi * np.asarray(day_of_year) / 365.25
    day = 2 * np.pi * np.asarray(time_of_day, dtype=np.float64) / 24.0
    return np.stack([np.sin(year), np.cos(year), np.sin(day), np.cos(day)], axis=-1)


def prepare_features(weather, time_data, koppen_row, normalize=True, stats=None):
    """featurePreprocessor.py:84-182 on plain arrays, statement by statement: ``weather`` [time, lat, lon, 12] (NaN
    allowed, modified in place like the reference does), ``time_data`` [time, 4], ``koppen_row`` [1, 8] torch tensor.
    Returns (combined f32 [time, nodes, 24], stats)."""
    weather_data = weather
    if np.isnan(weather_data).sum() > 0:  # :97-109
        for i in range(weather_data.shape[-1]):
            var_data = weather_data[..., i]
            var_mean = np.nanmean(var_data) if not np.all(np.isnan(var_data)) else np.nan
            if np.isnan(var_mean):
                var_mean = 0.0
            weather_data[..., i] = np.nan_to_num(var_data, nan=var_mean)
    num_time, num_lat, num_lon, num_weather = weather_data.shape  # :117-122
    num_nodes = num_lat * num_lon
    weather_features = weather_data.reshape(num_time, num_nodes, num_weather)
    if normalize:  # :125-146
        if stats is not None:
            mean, std = np.array(stats["mean"]), np.array(stats["std"])
        else:
            mean = weather_features.mean(axis=(0, 1))
            std = weather_features.std(axis=(0, 1)) + 1e-8
            if np.any(np.isnan(mean)) or np.any(np.isnan(std)):
                mean, std = np.nan_to_num(mean, nan=0.0), np.nan_to_num(std, nan=1.0)
            stats = {"mean": mean, "std": std}
        weather_features = (weather_features - mean) / std
    elif stats is None:
        stats = {}
    weather_tensor = torch.tensor(weather_features, dtype=torch.float32)  # :161-176
    time_tensor = torch.tensor(np.tile(time_data[:, np.newaxis, :], (1, num_nodes, 1)), dtype=torch.float32)
    koppen_expanded = koppen_row.detach().cpu().unsqueeze(0).expand(num_time, num_nodes, -1)
    combined = torch.cat([weather_tensor, time_tensor, koppen_expanded], dim=-1)
    if torch.isnan(combined).any():  # :178-180
        combined = torch.nan_to_num(combined, nan=0.0)
    return combined, stats
