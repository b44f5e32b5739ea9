"""De-normalisation helpers and per-variable forecast metrics: the part of
featurePreprocessor.py:187-239 and validate_hybrid_v5.py:338-358 that consumes the model's output.  The ingest half of
the reference's featurePreprocessor (NetCDF -> [time, N, 24]) is out of scope; only its output layout is contract."""
from __future__ import annotations

import numpy as np
import torch

# channel order of the 12 weather variables in features[..., :12] (featurePreprocessor.py:42-55)
WEATHER_VARS = ["u10", "v10", "t2m", "d2m", "sp", "tp", "u100", "v100", "str", "hcc", "lcc", "e"]


def denormalize_predictions(predictions, stats, target_var_idx=2):
    """featurePreprocessor.py:185-213 -- ``predictions * std[idx] + mean[idx]`` (default idx 2 = t2m); returned unchanged
    when ``stats`` carries no 'mean' / 'std'."""
    if "mean" in stats and "std" in stats:
        mean, std = stats["mean"][target_var_idx], stats["std"][target_var_idx]
        if isinstance(predictions, torch.Tensor):
            mean = torch.tensor(mean, dtype=predictions.dtype, device=predictions.device)
            std = torch.tensor(std, dtype=predictions.dtype, device=predictions.device)
        return predictions * std + mean
    return predictions


def denormalize_all_predictions(predictions, stats):
    """featurePreprocessor.py:216-239 -- all 12 variables; ``predictions`` [samples, 12] or [12]."""
    mean, std = stats["mean"], stats["std"]
    if isinstance(predictions, torch.Tensor):
        mean = torch.as_tensor(np.asarray(mean), dtype=predictions.dtype, device=predictions.device)
        std = torch.as_tensor(np.asarray(std), dtype=predictions.dtype, device=predictions.device)
    if predictions.ndim == 1:
        return predictions * std + mean
    return predictions * std[None, :] + mean[None, :]


def forecast_metrics(y_pred, y_true, stats, num_nodes, horizon, var_names=None, exclude_from_average=("sp",)):
    """Per-variable MSE / MAE of node-averaged, de-normalised forecasts (validate_hybrid_v5.py:220-223,338-358).

    ``y_pred`` / ``y_true``: [horizon * num_nodes, 12] in the reference's row order (they are reshaped to
    [horizon, num_nodes, 12] and averaged over nodes exactly as the reference does); the first six variables are
    reported and surface pressure is left out of ``average_mse``."""
    names = list(WEATHER_VARS[:6] if var_names is None else var_names)
    to_np = lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    mean, std = np.asarray(stats["mean"]), np.asarray(stats["std"])
    yp = to_np(y_pred).reshape(horizon, num_nodes, 12).mean(axis=1)
    yt = to_np(y_true).reshape(horizon, num_nodes, 12).mean(axis=1)
    out, total, count = {}, 0.0, 0
    for v, name in enumerate(names):
        if v >= yt.shape[1]:
            continue
        t, p = yt[:, v] * std[v] + mean[v], yp[:, v] * std[v] + mean[v]
        mse, mae = float(np.mean((p - t) ** 2)), float(np.mean(np.abs(p - t)))
        out[name] = {"mse": mse, "mae": mae}
        if name not in exclude_from_average:
            total, count = total + mse, count + 1
    out["average_mse"] = total / count if count else 0
    return out
