"""torch.autograd.Functions over the C-ABI launchers (single parameter set, G = 1).

These back the drop-in ``nn.Module`` classes (model.py, hybrid_model.py) so that user code
written against the reference -- ``out = model(batch.x, batch.edge_index); loss.backward()`` --
runs unchanged on the CUDA kernels.  The task-batched MAML path does not go through autograd
at all (engine.py / train_hybrid_maml_v5.py).
"""
from __future__ import annotations

import torch

from . import _lib
from .graph import RegionGraph


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class GCNConvReLU(torch.autograd.Function):
    """Y = [relu]((A_hat X) W^T + b) for one or more windows sharing a graph (wf_gcn_layer_fwd/bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, graph: RegionGraph, relu: bool):
        _lib.require_cuda(x, weight, bias)
        x, weight, bias = _f32c(x), _f32c(weight), _f32c(bias)
        rows, cin = x.shape
        cout = weight.shape[0]
        if rows % graph.R != 0:
            raise ValueError(f"x has {rows} rows, graph was normalised over {graph.R}")
        bw = rows // graph.R
        y = torch.empty(rows, cout, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("wf_gcn_layer_fwd", _lib.ptr(x), cin, graph.R * cin, None, _lib.ptr(weight), _lib.ptr(bias), 0, 0,
                      _lib.ptr(graph.rowptr), _lib.ptr(graph.col), _lib.ptr(graph.val), 0, 0, graph.R, cin, cout, 1,
                      bw, int(relu), _lib.ptr(y), _lib.stream_ptr())
        ctx.graph, ctx.relu, ctx.bw = graph, bool(relu), bw
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        graph, bw = ctx.graph, ctx.bw
        rows, cin = x.shape
        cout = weight.shape[0]
        dy = _f32c(dy).clone()  # overwritten with dY * (Y > 0)
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(weight) if need_w else None
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if need_b else None
        nbytes = int(_lib.query("wf_gcn_layer_bwd_workspace_bytes", graph.R, cin, cout, 1, bw))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("wf_gcn_layer_bwd", _lib.ptr(x), cin, graph.R * cin, None, _lib.ptr(y), _lib.ptr(dy),
                      _lib.ptr(weight), 0, _lib.ptr(graph.rowptr), _lib.ptr(graph.col), _lib.ptr(graph.val),
                      _lib.ptr(graph.rowptr_t), _lib.ptr(graph.col_t), _lib.ptr(graph.val_t), 0, 0, graph.R, cin, cout,
                      1, bw, int(ctx.relu), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), 0, 0, _lib.ptr(ws), nbytes,
                      _lib.stream_ptr())
        return dx, dw, db, None, None


class LSTMHead(torch.autograd.Function):
    """feats [T*N, F] -> predictions [N, O]: multi-layer LSTM over every node + Linear head.

    ``flat`` is the trainable flat buffer (engine.trainable_layout order); its gradient comes
    back flat as well.  Features get no gradient: the reference detaches them
    (hybrid_model.py:63).
    """

    @staticmethod
    def forward(ctx, feats, flat, dims, bw):
        _lib.require_cuda(feats, flat)
        feats, flat = _f32c(feats), _f32c(flat)
        d = dims
        rows = bw * d.R
        if feats.shape != (rows, d.hidden):
            raise ValueError(f"features {tuple(feats.shape)} do not match [{rows}, {d.hidden}]")
        Ls, L = d.lstm_layers, d.lstm_hidden
        f32 = dict(dtype=torch.float32, device=feats.device)
        gates = torch.empty(Ls, rows, 4 * L, **f32)
        h = torch.empty(Ls, rows, L, **f32)
        c = torch.empty(Ls, rows, L, **f32)
        pred = torch.empty(bw * d.num_nodes, d.O, **f32)
        with torch.cuda.device(feats.device):
            st = _lib.stream_ptr()
            _lib.call("wf_lstm_fwd", _lib.ptr(feats), _lib.ptr(flat), 0, Ls, d.hidden, L, d.O, d.window, d.num_nodes, 1,
                      bw, _lib.ptr(gates), _lib.ptr(h), _lib.ptr(c), st)
            _lib.call("wf_head_fwd", _lib.ptr(h[Ls - 1]), _lib.ptr(flat), 0, Ls, d.hidden, L, d.O, d.window,
                      d.num_nodes, 1, bw, _lib.ptr(pred), st)
        ctx.dims, ctx.bw = d, bw
        ctx.save_for_backward(feats, flat, gates, h, c)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        feats, flat, gates, h, c = ctx.saved_tensors
        d, bw = ctx.dims, ctx.bw
        Ls, L = d.lstm_layers, d.lstm_hidden
        dpred = _f32c(dpred)
        f32 = dict(dtype=torch.float32, device=feats.device)
        dlast = torch.empty(bw * d.num_nodes, L, **f32)
        grads = torch.zeros(flat.numel(), **f32)
        nbytes = int(max(_lib.query("wf_lstm_bwd_workspace_bytes", Ls, d.hidden, L, d.window, d.num_nodes, 1, bw),
                         _lib.query("wf_head_workspace_bytes", L, d.O, d.num_nodes, 1, bw)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=feats.device)
        gates = gates.clone()  # BPTT overwrites the activations; keep the graph re-usable
        with torch.cuda.device(feats.device):
            st = _lib.stream_ptr()
            _lib.call("wf_head_bwd", _lib.ptr(dpred), _lib.ptr(h[Ls - 1]), _lib.ptr(flat), 0, Ls, d.hidden, L, d.O,
                      d.window, d.num_nodes, 1, bw, _lib.ptr(dlast), _lib.ptr(grads), grads.numel(), _lib.ptr(ws),
                      nbytes, st)
            _lib.call("wf_lstm_bwd", _lib.ptr(feats), _lib.ptr(flat), 0, Ls, d.hidden, L, d.O, d.window, d.num_nodes, 1,
                      bw, _lib.ptr(gates), _lib.ptr(h), _lib.ptr(c), _lib.ptr(dlast), _lib.ptr(grads), grads.numel(),
                      _lib.ptr(ws), nbytes, st)
        return None, grads, None, None


def gcn_conv(x, weight, bias, graph, relu=False):
    return GCNConvReLU.apply(x, weight, bias, graph, relu)


def lstm_head(feats, flat, dims, bw=1):
    return LSTMHead.apply(feats, flat, dims, bw)
