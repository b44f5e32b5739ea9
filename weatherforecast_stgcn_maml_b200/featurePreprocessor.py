"""Drop-in for the reference's ``featurePreprocessor``.

* ``prepare_model_input`` (featurePreprocessor.py:66-182): NaN fill with the per-variable nanmean, z-score over
  (time, nodes), time features and the Koppen embedding row concatenated to ``[time, nodes, 24]`` -- here two streaming
  CUDA passes (``wf_feature_stats``, ``wf_assemble_features``, csrc/wf_features.cu) whose result STAYS in HBM, where the
  windowing (dataset.py) addresses it in place.  NetCDF ingest (dataLoader.py) stays out of scope: ``ds`` is anything
  indexable by variable name whose items have ``.values`` (an xarray Dataset, a dict of small wrappers, ...).
* ``denormalize_predictions`` / ``denormalize_all_predictions`` (:185-239) and ``forecast_metrics``
  (validate_hybrid_v5.py:338-358): the part that consumes the model's output.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

# channel order of the 12 weather variables in features[..., :12] (featurePreprocessor.py:42-55)
WEATHER_VARS = ["u10", "v10", "t2m", "d2m", "sp", "tp", "u100", "v100", "str", "hcc", "lcc", "e"]
# featurePreprocessor.py:59-64 (embed_utils.add_time_embeddings writes them)
TIME_VARS = ["year_progress_sin", "year_progress_cos", "day_progress_sin", "day_progress_cos"]


def feature_stats(weather):
    """Device pass over ``weather`` (CUDA f32 ``[..., 12]``): dict of numpy arrays ``fill`` (f32 nanmean per variable, 0
    where a variable is all NaN), ``mean`` / ``std`` (f64; of the FILLED array over every leading axis, population std,
    no epsilon) and ``nan_count`` (featurePreprocessor.py:104-109, :133-136)."""
    _lib.require_cuda(weather)
    if weather.dtype != torch.float32 or weather.shape[-1] != 12 or not weather.is_contiguous():
        raise ValueError(f"weather must be contiguous f32 [..., 12], got {weather.dtype} {tuple(weather.shape)}")
    rows = weather.numel() // 12
    dev = weather.device
    fill = torch.empty(12, dtype=torch.float32, device=dev)
    mean = torch.empty(12, dtype=torch.float64, device=dev)
    std = torch.empty(12, dtype=torch.float64, device=dev)
    nans = torch.empty(12, dtype=torch.int64, device=dev)
    nbytes = _lib.query("wf_feature_stats_workspace_bytes", rows)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("wf_feature_stats", _lib.ptr(weather), rows, _lib.ptr(fill), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(nans),
                  _lib.ptr(ws), nbytes, _lib.stream_ptr())
    return {"fill": fill.cpu().numpy(), "mean": mean.cpu().numpy(), "std": std.cpu().numpy(), "nan_count": nans.cpu().numpy()}


def assemble_features(weather, time_features, koppen_row, fill=None, mean=None, std=None, f64_arith=False):
    """``[time, N, 24]`` CUDA f32 from ``weather [time, N, 12]`` (CUDA f32, NaN allowed), ``time_features [time, 4]`` and
    the 8-value Koppen embedding row; ``mean``/``std`` None = no normalisation (featurePreprocessor.py:146, :164-180)."""
    _lib.require_cuda(weather)
    T, N = int(weather.shape[0]), int(weather.shape[1])
    dev = weather.device
    tf = torch.as_tensor(np.ascontiguousarray(time_features, dtype=np.float32)).to(dev) if not isinstance(
        time_features, torch.Tensor) else time_features.to(device=dev, dtype=torch.float32).contiguous()
    if tuple(tf.shape) != (T, 4):
        raise ValueError(f"time_features must be [{T}, 4], got {tuple(tf.shape)}")
    out = torch.empty(T, N, 24, dtype=torch.float32, device=dev)
    h_fill = np.zeros(12, dtype=np.float32) if fill is None else np.ascontiguousarray(fill, dtype=np.float32)
    normalize = mean is not None
    h_mean = np.ascontiguousarray(mean if normalize else np.zeros(12), dtype=np.float64)
    h_std = np.ascontiguousarray(std if normalize else np.ones(12), dtype=np.float64)
    h_kop = np.ascontiguousarray(koppen_row, dtype=np.float32).reshape(-1)
    if h_kop.size != 8 or h_mean.size != 12 or h_std.size != 12 or h_fill.size != 12:
        raise ValueError("fill / mean / std need 12 values, the Koppen row 8")
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    with torch.cuda.device(dev):
        _lib.call("wf_assemble_features", _lib.ptr(weather), T, N, hp(h_fill), hp(h_mean), hp(h_std), int(normalize),
                  int(bool(f64_arith)), _lib.ptr(tf), hp(h_kop), _lib.ptr(out), _lib.stream_ptr())
    return out


def prepare_model_input(ds, koppen_code, koppen_embed_layer, normalize=True, stats=None, device="cuda"):
    """featurePreprocessor.py:66-182 with the same arguments, return values and arithmetic; ``features`` is a CUDA
    tensor ``[time, nodes, 24]`` (the reference returns a CPU tensor and copies one window per step to the device).

    The statistics the reference derives itself are f32 (numpy reductions over f32 data, ``std + 1e-8``) and are applied
    in f32; statistics passed in are lists -> f64 arrays and are applied in f64 -- both reproduced."""
    weather = np.stack([np.asarray(ds[var].values) for var in WEATHER_VARS], axis=-1)  # [time, lat, lon, 12]
    time_data = np.stack([np.asarray(ds[var].values) for var in TIME_VARS], axis=-1)  # [time, 4]
    num_time = weather.shape[0]
    num_nodes = int(np.prod(weather.shape[1:-1]))
    w = torch.as_tensor(np.ascontiguousarray(weather, dtype=np.float32)).reshape(num_time, num_nodes, 12).to(device)
    st = feature_stats(w)
    fill = st["fill"] if int(st["nan_count"].sum()) > 0 else None
    f64_arith = False
    mean = std = None
    if normalize:
        if stats is not None:
            mean, std = np.array(stats["mean"]), np.array(stats["std"])
            f64_arith = mean.dtype == np.float64 or std.dtype == np.float64
        else:
            mean = st["mean"].astype(np.float32)
            std = st["std"].astype(np.float32) + np.float32(1e-8)
            if np.any(np.isnan(mean)) or np.any(np.isnan(std)):
                mean, std = np.nan_to_num(mean, nan=0.0), np.nan_to_num(std, nan=1.0)
            stats = {"mean": mean, "std": std}
    elif stats is None:
        stats = {}
    kdev = next(koppen_embed_layer.parameters()).device
    with torch.no_grad():
        row = koppen_embed_layer(torch.tensor([koppen_code], dtype=torch.long, device=kdev)).detach().cpu().numpy()
    feats = assemble_features(w, time_data.astype(np.float32), row, fill=fill, mean=mean, std=std, f64_arith=f64_arith)
    return feats, stats


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
