"""Drop-in for the reference's ``hybrid_model.HybridSTGCN_LSTM`` (hybrid_model.py:6-134).

Same constructor, methods and ``state_dict`` (28 tensors, 834,752 parameters for the v5
configuration).  ``forward(x, edge_index)`` reproduces the reference semantics:

* the four GCN layers run under ``torch.no_grad()`` unconditionally (hybrid_model.py:63), so
  no gradient reaches ``base_stgcn`` even when it is "unfrozen" (SURVEY.md D4);
* the per-node Python loop over ``nn.LSTM`` (hybrid_model.py:93-105, one call per node) is one
  batched recurrence over all nodes (functional.LSTMHead -> wf_lstm_fwd / wf_lstm_bwd);
* predictions come back as ``[N*H, out]`` with row = node*H + h (hybrid_model.py:114-115).

Dropout: in train mode the three sites of the reference (after GCN layers 1-3 at
hybrid_model.py:67-73, between the LSTM layers :47, on the head input :108) are fused into the
kernels as counter-based masks (csrc/wf_rng.cuh) that the backward pass regenerates; torch's
generator stream cannot be matched by a batched kernel, the distribution is.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as WF
from .engine import V5Dims, model_dropout, trainable_layout
from .model import STGCN  # noqa: F401  (same import as the reference module)


class LSTMParameters(nn.Module):
    """Parameter holder with ``nn.LSTM``'s names, shapes, init and ``flatten_parameters()``."""

    def __init__(self, input_size, hidden_size, num_layers, dropout=0.0):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.dropout = input_size, hidden_size, num_layers, dropout
        self.batch_first, self.bidirectional = True, False
        for l in range(num_layers):
            kin = input_size if l == 0 else hidden_size
            self.register_parameter(f"weight_ih_l{l}", nn.Parameter(torch.empty(4 * hidden_size, kin)))
            self.register_parameter(f"weight_hh_l{l}", nn.Parameter(torch.empty(4 * hidden_size, hidden_size)))
            self.register_parameter(f"bias_ih_l{l}", nn.Parameter(torch.empty(4 * hidden_size)))
            self.register_parameter(f"bias_hh_l{l}", nn.Parameter(torch.empty(4 * hidden_size)))
        self.reset_parameters()

    def reset_parameters(self):
        a = 1.0 / math.sqrt(self.hidden_size)
        with torch.no_grad():
            for p in self.parameters():
                p.uniform_(-a, a)

    def flatten_parameters(self):
        """No-op: the kernels read a flat copy assembled per call (train_hybrid_maml_v5.py:212)."""
        return None


class HybridSTGCN_LSTM(nn.Module):
    def __init__(self, base_stgcn, lstm_hidden_size=64, lstm_num_layers=2, lstm_dropout=0.2, out_channels=12,
                 forecast_horizon=8, freeze_base=True):
        super().__init__()
        self.forecast_horizon = forecast_horizon
        self.out_channels = out_channels
        self.lstm_hidden_size = lstm_hidden_size
        self.base_stgcn = base_stgcn
        if freeze_base:
            for param in self.base_stgcn.parameters():
                param.requires_grad = False
        base_hidden_channels = self.base_stgcn.conv1.out_channels
        self.lstm = LSTMParameters(base_hidden_channels, lstm_hidden_size, lstm_num_layers,
                                   dropout=lstm_dropout if lstm_num_layers > 1 else 0.0)
        self.output_layer = nn.Linear(lstm_hidden_size, out_channels * forecast_horizon)
        self.dropout = nn.Dropout(lstm_dropout)

    # -- helpers -------------------------------------------------------------------------
    def dims(self, num_nodes):
        b = self.base_stgcn
        return V5Dims(num_nodes=num_nodes, window=b.window_size, horizon=self.forecast_horizon,
                      in_channels=b.conv1.in_channels, hidden=b.conv1.out_channels,
                      lstm_hidden=self.lstm_hidden_size, lstm_layers=self.lstm.num_layers,
                      out_channels=self.out_channels, num_weather=self.out_channels)

    def flat_trainable(self, dims):
        sd = dict(self.named_parameters())
        return torch.cat([sd[name].reshape(-1) for name, _, _ in trainable_layout(dims)])

    # -- reference API -------------------------------------------------------------------
    def extract_base_features(self, x, edge_index):
        with torch.no_grad():
            b = self.base_stgcn
            p = float(b.dropout.p) if b.dropout.training else 0.0  # nn.Dropout follows its own .training flag
            h = b.conv1(x, edge_index, _fuse_relu=True, _dropout=p, _site=0)
            h = b.conv2(h, edge_index, _fuse_relu=True, _dropout=p, _site=1)
            h = b.conv3(h, edge_index, _fuse_relu=True, _dropout=p, _site=2)
            h = b.conv4(h, edge_index, _fuse_relu=True)  # no final dropout (hybrid_model.py:76)
        return h

    def forward(self, x, edge_index):
        base_features = self.extract_base_features(x, edge_index)
        window = self.base_stgcn.window_size
        num_nodes = base_features.shape[0] // window
        dims = self.dims(num_nodes)
        _, p_lstm, p_head = model_dropout(self)
        drop = (p_lstm if self.lstm.training else 0.0, p_head if self.dropout.training else 0.0)
        pred = WF.lstm_head(base_features, self.flat_trainable(dims), dims, 1, dropout=drop)  # [N, H*out]
        return pred.view(num_nodes, self.forecast_horizon, self.out_channels).reshape(-1, self.out_channels)

    def get_trainable_parameters(self):
        trainable_params = []
        trainable_params.extend(self.lstm.parameters())
        trainable_params.extend(self.output_layer.parameters())
        return trainable_params

    def freeze_base_model(self):
        for param in self.base_stgcn.parameters():
            param.requires_grad = False

    def unfreeze_base_model(self):
        for param in self.base_stgcn.parameters():
            param.requires_grad = True
