"""Drop-in for the reference's ``model`` module: ``GCNConv`` and ``STGCN`` (model.py:7-52).

Same constructor signature, attribute names (``conv1..conv4``, ``dropout``, ``output_layer``,
``window_size``), ``forward(x [T*N, C], edge_index [2, E]) -> [H*N, out]`` and ``state_dict``
keys (``convN.bias``, ``convN.lin.weight``, ``output_layer.{weight,bias}``) as the reference
class over PyG >= 2.0.  The four convolutions run on the fused CUDA GCN layer
(functional.GCNConvReLU: tensor cores where the widths allow, ReLU and train-mode dropout in the
epilogue); the graph normalisation PyG recomputes on every call is cached per ``edge_index``
tensor (graph.graph_for); the Linear head of ``STGCN.forward`` is the C-ABI head kernel
(functional.LinearRows), not cuBLAS.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as WF
from .graph import graph_for


class _Lin(nn.Module):
    """Bias-free linear holder -> state_dict key ``<conv>.lin.weight`` (PyG ``Linear``)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))  # PyG 'glorot'
        with torch.no_grad():
            self.weight.uniform_(-a, a)


class GCNConv(nn.Module):
    """PyG ``GCNConv(in, out)`` with its defaults (add_self_loops, normalize, bias, sum aggregation,
    source -> target flow): ``out[i] = sum_{e: dst_e = i} w_e (x W^T)[src_e] + b``."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.bias = nn.Parameter(torch.zeros(out_channels))  # registered before ``lin`` like PyG
        self.lin = _Lin(in_channels, out_channels)
        self._w16 = None  # ((data_ptr, version, device), (hi, lo)): fp16 operand halves of lin.weight

    def reset_parameters(self):
        self.lin.reset_parameters()
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x, edge_index, _fuse_relu=False, _dropout=0.0, _site=0):
        """``_fuse_relu`` / ``_dropout`` / ``_site``: ReLU and nn.Dropout(p) (mask site = layer index) fused into the
        layer's epilogue -- what STGCN.forward and HybridSTGCN_LSTM.extract_base_features do right after the conv."""
        graph = graph_for(edge_index, x.shape[0], x.device)
        w16 = None
        w = self.lin.weight
        if w.is_cuda and WF.precision() == "tf32x3" and self.out_channels % 128 == 0 and self.in_channels % 8 == 0:
            # tensor-core operand halves of the weight, recomputed whenever it changes (optimiser steps and
            # load_state_dict bump _version; .to() / .cuda() move the storage)
            key = (w.data_ptr(), w._version, w.device)
            if self._w16 is None or self._w16[0] != key:
                self._w16 = (key, WF.split_weight16(w.detach()))
            w16 = self._w16[1]
        return WF.gcn_conv(x, w, self.bias, graph, relu=_fuse_relu, p_drop=_dropout, site=_site, w16=w16)

    def __deepcopy__(self, memo):
        # the operand cache belongs to THIS module's storage: a copy starts without one
        import copy

        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_w16" else copy.deepcopy(v, memo)
        return new


class STGCN(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels=12, window_size=6, forecast_horizon=1,
                 dropout_rate=0.3):
        super().__init__()
        self.window_size = window_size
        self.out_channels = out_channels
        self.forecast_horizon = forecast_horizon
        self.dropout_rate = dropout_rate
        self.conv1 = GCNConv(in_channels, hidden_channels)
        self.conv2 = GCNConv(hidden_channels, hidden_channels)
        self.conv3 = GCNConv(hidden_channels, hidden_channels)
        self.conv4 = GCNConv(hidden_channels, hidden_channels)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.output_layer = nn.Linear(hidden_channels, out_channels * forecast_horizon)

    def _p(self):
        """Probability of ``self.dropout`` (an nn.Dropout kept for attribute / state compatibility) when it is active."""
        return float(self.dropout.p) if self.training else 0.0

    def forward(self, x, edge_index):
        # conv -> relu -> dropout, four times (model.py:31-42), each as ONE fused layer launch
        p = self._p()
        for i, conv in enumerate((self.conv1, self.conv2, self.conv3, self.conv4)):
            x = conv(x, edge_index, _fuse_relu=True, _dropout=p, _site=i)
        num_nodes = x.shape[0] // self.window_size
        x = x[-num_nodes:]  # last time slice (model.py:45-48)
        x = WF.linear_rows(x, self.output_layer.weight, self.output_layer.bias)
        x = x.view(num_nodes, self.forecast_horizon, self.out_channels)
        return x.reshape(-1, self.out_channels)
