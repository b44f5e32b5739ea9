"""Regional fine-tuning loop of the reference's ``adapt_hybrid_v5`` (adapt_hybrid_v5.py:152-231).

The reference's ``adaptModel`` first reads a checkpoint and multi-year NetCDF files from
hard-coded paths (adapt_hybrid_v5.py:22,84,132) -- disk I/O that is out of scope.  What is on
the hot path, and reproduced here on the device, is everything after the data exists:

* split: first ``min(1200, len)`` windows, 80/20 chronological (:152-159);
* 15 epochs of batch-1 steps in shuffled order: forward, MSE, backward,
  clip_grad_norm_(1.0), Adam with coupled L2 chosen by region name (:185-203,
  adaptive_scheduler.py:68-95), lr updated per epoch by ClimateAwareLRScheduler (:208);
* eval-mode mean MSE over the validation windows (:216-231);
* the adapted checkpoint dict with the reference's keys (:240-257).

Each step is one CUDA-graph replay on an engine with G = 1, Bw = 1; the window is selected by
copying its offsets into a device slot, so the features never leave HBM.  Sequential batch-1
SGD over one region does not shard: across GPUs this is "replicas only" (one region per GPU,
as main.py:30 iterates regions independently).
"""
from __future__ import annotations

import torch

from . import _lib
from .adaptive_scheduler import ClimateAwareLRScheduler, climate_hyperparameters
from .engine import (REFERENCE_DROPOUT, AdamState, HybridEngine, V5Dims, flatten_trainable, gcn_weights_from_state_dict,
                     raise_on_error_code, unflatten_trainable)
from .graph import RegionGraph

EPOCHS = 15  # adapt_hybrid_v5.py:185


class FineTuner:
    def __init__(self, state_dict, features, edge_index, dims: V5Dims, device="cuda", region_name="",
                 base_lr=0.0006, max_samples=1200, train_frac=0.8, use_cuda_graph=True, val_batch=16,
                 dropout=REFERENCE_DROPOUT, seed=0):
        """``dropout`` = (p_gcn, p_lstm, p_head): the reference fine-tunes in ``.train()`` mode (adapt_hybrid_v5.py:168)
        on a model rebuilt with ``dropout_rate=0.2`` and the checkpoint's ``lstm_dropout`` (:99-117); validation runs in
        ``.eval()`` mode (:214).  Pass ``(0, 0, 0)`` for the deterministic parity configuration."""
        self.dims, self.device = dims, torch.device(device)
        d = dims
        self.features = features.to(self.device, torch.float32).contiguous()
        n_windows = self.features.shape[0] - d.window - d.horizon  # dataset.py:25
        self.max_samples = min(max_samples, n_windows)
        self.train_size = int(train_frac * self.max_samples)
        self.val_idx = list(range(self.train_size, self.max_samples))
        self.graph_csr = RegionGraph(edge_index, d.R, self.device)
        self.sd = {k: v.detach().clone() for k, v in state_dict.items()}
        self.theta = flatten_trainable(self.sd, d, self.device)
        self.gcn_w = gcn_weights_from_state_dict(self.sd, self.device)
        self.engine = HybridEngine(d, 1, 1, self.device, dropout=dropout, seed=seed)
        self.P = self.engine.P
        lr, wd = climate_hyperparameters(region_name, base_lr)
        self.initial_lr = lr
        self.adam = AdamState(self.P, self.device, lr, weight_decay=wd, decoupled=False)
        self.scheduler = ClimateAwareLRScheduler(self.adam, region_name, lr)
        per_step = d.num_nodes * d.in_channels
        idx = torch.arange(self.max_samples, dtype=torch.long)
        self.x_table = (idx * per_step).to(self.device)
        self.t_table = ((idx + d.window + 1) * per_step).to(self.device)
        self.cur_x = torch.zeros(1, dtype=torch.long, device=self.device)
        self.cur_t = torch.zeros(1, dtype=torch.long, device=self.device)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.err_seen = torch.zeros(1, dtype=torch.int32, device=self.device)  # kernel error flags, read with the loss
        self.use_graph, self.graph = bool(use_cuda_graph), None
        self.val_batch = val_batch
        self._val_engine = None
        self.steps_done = 0

    def _body(self):
        e, d = self.engine, self.dims
        C = d.in_channels
        e.forward_backward(self.features, C, 0, self.cur_x, self.gcn_w, self.graph_csr, self.theta, 0,
                           feat=self.features, tgt_off=self.cur_t, feat_ld=C, grad_scale=1.0)
        self.loss_sum += e.loss
        self.err_seen.copy_(torch.maximum(self.err_seen, e.err))

    def step(self, window_index):
        """One batch-1 training step on window ``window_index`` (adapt_hybrid_v5.py:193-201)."""
        self.cur_x.copy_(self.x_table[window_index:window_index + 1])
        self.cur_t.copy_(self.t_table[window_index:window_index + 1])
        if self.use_graph:
            if self.graph is None:
                keep = self.loss_sum.clone()
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._body()
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                self.loss_sum.copy_(keep)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._body()
            self.graph.replay()
        else:
            self._body()
        self.adam.step(self.theta, self.engine.grads.view(-1), max_norm=1.0)
        self.steps_done += 1

    def train_epoch(self, order=None):
        """One pass over the training windows; ``order`` defaults to a fresh ``torch.randperm``
        (DataLoader(shuffle=True), adapt_hybrid_v5.py:182).  Returns the mean training loss."""
        if order is None:
            order = torch.randperm(self.train_size).tolist()
        self.loss_sum.zero_()
        for i in order:
            self.step(int(i))
        total = float(self.loss_sum.item())
        self.check()
        return total / max(1, len(order))

    def check(self):
        """Raise if a kernel flagged an error since the last check (read at points that synchronise anyway)."""
        code = max(int(self.err_seen.item()), int(self.engine.err.item()))
        if self._val_engine is not None:
            code = max(code, int(self._val_engine.err.item()))
        raise_on_error_code(code)

    def fit(self, epochs=EPOCHS, orders=None, verbose=False):
        history = []
        for ep in range(epochs):
            avg = self.train_epoch(None if orders is None else orders[ep])
            lr = self.scheduler.step(avg)  # adapt_hybrid_v5.py:208
            history.append((avg, lr))
            if verbose:
                print(f"Epoch {ep + 1}/{epochs}: Loss = {avg:.6f}, LR = {lr:.6f}")
        return history

    @torch.no_grad()
    def validate(self, indices=None):
        """Mean per-window MSE over the validation windows in eval mode (adapt_hybrid_v5.py:216-231)."""
        d = self.dims
        indices = self.val_idx if indices is None else list(indices)
        if not indices:
            return float("nan")
        total = torch.zeros(1, dtype=torch.float32, device=self.device)
        C = d.in_channels
        pos = 0
        while pos < len(indices):
            chunk = indices[pos:pos + self.val_batch]
            pos += len(chunk)
            if self._val_engine is None or self._val_engine.Bw != len(chunk):
                self._val_engine = HybridEngine(d, 1, len(chunk), self.device, training=False)  # eval mode: no dropout
            e = self._val_engine
            sel = torch.tensor(chunk, dtype=torch.long, device=self.device)
            e.gcn_forward(self.features, C, 0, self.x_table[sel].contiguous(), self.gcn_w, self.graph_csr)
            e.lstm_head_forward(self.theta, 0)
            e.mse(feat=self.features, tgt_off=self.t_table[sel].contiguous(), feat_ld=C, want_grad=False)
            total += e.loss.sum()
        out = float(total.item()) / len(indices)
        self.check()
        return out

    def state_dict(self):
        self.check()
        out = {k: v.clone() for k, v in self.sd.items()}
        for name, t in unflatten_trainable(self.theta.detach().cpu(), self.dims).items():
            out[name] = t.clone()
        return out


def adapt_region(checkpoint, features, edge_index, region_coords, region_name, stats=None, device="cuda",
                 epochs=EPOCHS, orders=None, verbose=True, dropout=None):
    """The compute part of ``adaptModel`` (adapt_hybrid_v5.py:84-257) on in-memory inputs.

    ``checkpoint`` is a meta-training checkpoint dict (keys of train_hybrid_maml_v5.py:311-335);
    returns the adapted checkpoint dict with the keys of adapt_hybrid_v5.py:240-257."""
    config, hybrid_config = checkpoint["config"], checkpoint["hybrid_config"]
    dims = V5Dims(num_nodes=features.shape[1], window=config["window_size"], horizon=config["forecast_horizon"],
                  in_channels=config["input_channels"], hidden=config["hidden_channels"],
                  lstm_hidden=hybrid_config["lstm_hidden_size"], lstm_layers=hybrid_config["lstm_num_layers"],
                  out_channels=config["output_channels"], num_weather=config["output_channels"])
    if dropout is None:  # adapt_hybrid_v5.py:99-117: STGCN(dropout_rate=0.2), lstm_dropout from the checkpoint
        p = float(hybrid_config.get("lstm_dropout", REFERENCE_DROPOUT[1]))
        dropout = (REFERENCE_DROPOUT[0], p, p)
    tuner = FineTuner(checkpoint["hybrid_model_state_dict"], features, edge_index, dims, device, region_name,
                      dropout=dropout)
    tuner.fit(epochs, orders, verbose)
    val = tuner.validate()
    sd = tuner.state_dict()
    return {
        "hybrid_model_state_dict": sd,
        "koppen_embed_state_dict": checkpoint["koppen_embed_state_dict"],
        "region": region_coords,
        "region_name": region_name,
        "climate_type": "Adapted_Region",
        "stats": stats,
        "config": config,
        "hybrid_config": hybrid_config,
        "model_version": "5.0",
        "adaptation_type": "v5_regional_adaptation_adaptive",
        "val_loss": val,
        "base_model_loss": checkpoint.get("meta_loss", "N/A"),
        "total_params": sum(v.numel() for v in sd.values()),
    }


MODEL_PATH = "./Out_Data/SavedModels/hybrid_maml_model_v5_best.pt"   # adapt_hybrid_v5.py:17
SAVE_DIR = "./Out_Data/AdaptedModels"                                # adapt_hybrid_v5.py:236


def adaptModel(region_coords, region_name, loader=None, model_path=None, save_dir=None, device="cuda", epochs=EPOCHS,
               orders=None, verbose=True, dropout=None):
    """Adapt Model V5 to a specific region -- the reference's entry point (adapt_hybrid_v5.py:65-271), same positional
    signature and return value (the path of the saved adapted checkpoint).

    What the reference does around the compute is disk I/O against hard-coded paths (``load_adaptation_data``: NetCDF
    files under ``E:/Study/...``, :22-62), which is out of scope; ``loader(region_coords)`` supplies the region's data
    instead and must return what that function returns: a dataset-like object exposing ``.latitude.values``,
    ``.longitude.values`` and, by name, the 12 ERA5 variables plus the four time features (``ds[name].values``; see
    featurePreprocessor.WEATHER_VARS / TIME_VARS) -- or, for data that is already assembled, a tuple
    ``(features [time, N, 24], edge_index [2, E], stats)``.  Everything else follows the reference line by line:
    checkpoint from ``MODEL_PATH`` (:84), kNN graph with k = 4 (:139), ``prepare_model_input(ds, 0, koppen_embed,
    normalize=True)`` (:140), 80/20 split of the first 1,200 windows, 15 epochs of batch-1 Adam fine-tuning in train mode,
    eval-mode validation (:152-231), checkpoint dict and file name of :236-257."""
    import os

    from .embed_utils import KoppenEmbedding
    from .featurePreprocessor import prepare_model_input
    from .graphBuilder import build_spatial_graph

    if loader is None:
        raise ValueError("adaptModel needs loader(region_coords): the reference's NetCDF reader (adapt_hybrid_v5.py:30-62, "
                         "hard-coded E:/ paths) is not part of this package")
    model_path = MODEL_PATH if model_path is None else model_path
    save_dir = SAVE_DIR if save_dir is None else save_dir
    if verbose:
        print("=" * 80)
        print(f"MODEL 5.0 REGIONAL ADAPTATION: {region_name}")
        print(f"Device: {device}  Region: {region_coords}  Base Model: {model_path}")
        print("=" * 80)
    checkpoint = torch.load(model_path, map_location="cpu", weights_only=False)
    data = loader(region_coords)
    if isinstance(data, tuple):
        features, edge_index, stats = data
    else:
        koppen_embed = KoppenEmbedding(embedding_dim=8)
        koppen_embed.load_state_dict(checkpoint["koppen_embed_state_dict"])
        edge_index, _, _ = build_spatial_graph(data, k_neighbors=4, device=device)
        features, stats = prepare_model_input(data, 0, koppen_embed.to(device), normalize=True, device=device)
    out = adapt_region(checkpoint, features, edge_index, region_coords, region_name, stats=stats, device=device,
                       epochs=epochs, orders=orders, verbose=verbose, dropout=dropout)
    os.makedirs(save_dir, exist_ok=True)
    save_path = os.path.join(save_dir, f"hybrid_v5_adapted_{region_name}_{region_coords}.pt")
    torch.save(out, save_path)
    if verbose:
        print(f"Final validation loss: {out['val_loss']:.6f}\nModel saved: {save_path}")
    return save_path
