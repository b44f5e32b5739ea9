"""Task-batched functional engine for the v5 hybrid model on flat device buffers.

One ``HybridEngine`` owns every activation / workspace buffer for a fixed batch shape
(G tasks x Bw windows x N nodes) and drives the C-ABI launchers of libwf_stgcn.so:

    GCN stack (no grad, hybrid_model.py:60-78)  -> wf_gcn_layer_fwd x4
    LSTM over all (task, window, node) sequences -> wf_lstm_fwd         (hybrid_model.py:93-105)
    head + per-window MSE + backward seed        -> wf_head_fwd, wf_mse_fwd_bwd, wf_head_bwd
    BPTT                                          -> wf_lstm_bwd        (loss.backward())
    clip + SGD per task                           -> wf_clip_sgd_step   (train_hybrid_maml_v5.py:135-139)

Nothing here allocates after construction, synchronises, or touches the host, so a whole MAML
inner step (or meta-step) can be captured in a CUDA graph.  Weights are *flat*: the 18 tensors
autograd reaches on the hybrid path, concatenated in state_dict order, one copy
per task (the replacement for ``copy.deepcopy(model)`` at train_hybrid_maml_v5.py:111).
"""
from __future__ import annotations

import functools
import os
from dataclasses import dataclass

import torch

from . import _lib
from .graph import RegionGraph, StackedGraphs


@dataclass(frozen=True)
class V5Dims:
    """Model/config constants (train_hybrid_maml_v5.py:31-38)."""
    num_nodes: int = 441
    window: int = 24
    horizon: int = 8
    in_channels: int = 24
    hidden: int = 256
    lstm_hidden: int = 128
    lstm_layers: int = 4
    out_channels: int = 12
    num_weather: int = 12

    @property
    def R(self):
        return self.window * self.num_nodes

    @property
    def O(self):
        return self.out_channels * self.horizon

    @property
    def P(self):
        return int(_lib.query("wf_param_count", self.lstm_layers, self.hidden, self.lstm_hidden, self.O))


def trainable_layout(dims: V5Dims):
    """[(state_dict key, shape, offset)] of the flat trainable buffer."""
    L, out, off = dims.lstm_hidden, [], 0
    for l in range(dims.lstm_layers):
        kin = dims.hidden if l == 0 else L
        for name, shape in ((f"lstm.weight_ih_l{l}", (4 * L, kin)), (f"lstm.weight_hh_l{l}", (4 * L, L)),
                            (f"lstm.bias_ih_l{l}", (4 * L,)), (f"lstm.bias_hh_l{l}", (4 * L,))):
            out.append((name, shape, off))
            off += int(torch.Size(shape).numel())
    for name, shape in (("output_layer.weight", (dims.O, L)), ("output_layer.bias", (dims.O,))):
        out.append((name, shape, off))
        off += int(torch.Size(shape).numel())
    return out


def flatten_trainable(sd, dims: V5Dims, device=None, prefix=""):
    parts = [sd[prefix + name].detach().reshape(-1).to(torch.float32) for name, _, _ in trainable_layout(dims)]
    flat = torch.cat(parts)
    return flat.to(device) if device is not None else flat


def unflatten_trainable(flat, dims: V5Dims):
    return {name: flat[off:off + int(torch.Size(shape).numel())].view(shape) for name, shape, off in trainable_layout(dims)}


def gcn_weights_from_state_dict(sd, device, prefix="base_stgcn."):
    """[(W [Cout, Cin], b [Cout])] x 4, contiguous on ``device``."""
    return [(sd[f"{prefix}conv{i}.lin.weight"].detach().to(device, torch.float32).contiguous(),
             sd[f"{prefix}conv{i}.bias"].detach().to(device, torch.float32).contiguous()) for i in range(1, 5)]


# Dropout probabilities of the reference's training configuration: (GCN, LSTM inter-layer, head input) =
# (STGCN dropout_rate, lstm_dropout, lstm_dropout) at train_hybrid_maml_v5.py:197,205.
REFERENCE_DROPOUT = (0.2, 0.2, 0.2)

ERR_MESSAGES = {41: "a GCN activation left the fp16 operand range (|x| >= 32768 or non-finite): normalise the input features "
                    "(prepare_model_input(normalize=True)) or build the engine with precision='fp32'"}


def model_dropout(hybrid_model):
    """(p_gcn, p_lstm, p_head) of a HybridSTGCN_LSTM as the reference applies them (hybrid_model.py:47,58,67-73,108)."""
    lstm = hybrid_model.lstm
    p_lstm = float(getattr(lstm, "dropout", 0.0)) if getattr(lstm, "num_layers", 1) > 1 else 0.0
    return (float(hybrid_model.base_stgcn.dropout.p), p_lstm, float(hybrid_model.dropout.p))


def _on_device(fn):
    """Run a launcher method with the engine's device current (streams and launches follow the current device)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


class HybridEngine:
    """precision="tf32x3": dense products on the tcgen05 tensor cores with the 3xTF32 split
    (FP32-class accuracy, csrc/wf_tc.cuh); used when the model shape allows it (GCN width a
    multiple of 128, LSTM hidden 128 -- the v5 configuration).  precision="fp32": every product
    on the exact FP32 CUDA-core kernels (also the route for other shapes)."""

    def __init__(self, dims: V5Dims, G: int, Bw: int, device="cuda", keep_gcn_activations=False, precision="tf32x3",
                 training=True, lstm_mode="persistent", dropout=(0.0, 0.0, 0.0), seed=0):
        """dropout = (p_gcn, p_lstm, p_head): train-mode nn.Dropout at the reference's three sites (after GCN layers
        1-3, between the LSTM layers, on the head input), as counter-based masks regenerated in the backward pass
        (csrc/wf_rng.cuh); active while ``self.stochastic`` is set (``train()`` / ``eval()``)."""
        self.dims, self.G, self.Bw = dims, int(G), int(Bw)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HybridEngine needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dropout = tuple(float(p) for p in dropout)
        if len(self.dropout) != 3 or any(not (0.0 <= p < 1.0) for p in self.dropout):
            raise ValueError("dropout must be three probabilities in [0, 1): (GCN, LSTM inter-layer, head)")
        if dims.lstm_layers < 2:
            self.dropout = (self.dropout[0], 0.0, self.dropout[2])  # nn.LSTM ignores dropout for a single layer
        self.stochastic = bool(training) and any(p > 0 for p in self.dropout)
        if precision not in ("tf32x3", "fp32"):
            raise ValueError("precision must be 'tf32x3' or 'fp32'")
        _lib.load()
        d = dims
        self.tc = precision == "tf32x3" and d.lstm_hidden == 128 and d.hidden % 128 == 0
        if lstm_mode not in ("persistent", "stepwise"):
            raise ValueError("lstm_mode must be 'persistent' or 'stepwise'")
        if self.tc and lstm_mode == "stepwise" and (self.dropout[1] > 0 or self.dropout[2] > 0):
            raise ValueError("lstm_mode='stepwise' (the 3xTF32 cross-check path) has no dropout; use 'persistent' or precision='fp32'")
        # persistent: one cluster launch per LSTM layer runs all T steps (csrc/wf_lstm_seq.cu);
        # stepwise: one tensor-core launch per (layer, step) (csrc/wf_tc_gemm.cu), kept as a cross-check
        self.seq = self.tc and lstm_mode == "persistent"
        self.training = bool(training)
        self.rows = self.G * self.Bw * d.R
        self.W = self.G * self.Bw
        f32 = dict(dtype=torch.float32, device=self.device)
        n_act = 4 if keep_gcn_activations else 2
        i16 = dict(dtype=torch.int16, device=self.device)
        self.keep_gcn = keep_gcn_activations
        Ls, L = d.lstm_layers, d.lstm_hidden
        # the GCN stack of the persistent path works on pre-split fp16 hi/lo planes (csrc/wf_gemm_ss.cu)
        self.gcn_ss = self.seq and d.in_channels % 8 == 0
        if self.gcn_ss:
            self.act16 = [torch.empty(2, self.rows, d.hidden, **i16) for _ in range(n_act)]
            self.featsb16 = torch.empty(2, self.rows, d.hidden, **i16) if self.training else None  # bf16 planes of the output
            self.x16 = torch.empty(2, self.rows, d.in_channels, **i16)   # split of the fp32 input windows (first layer)
            self.side16 = None                                            # aggregated leading rows, sized by the graph
            self.act = None
        else:
            self.act = [torch.empty(self.rows, d.hidden, **f32) for _ in range(n_act)]
        if self.seq:  # gates / cell state in the tile-blocked TB4 layout (padded to 128-node tiles)
            self.gates = torch.empty(Ls, int(_lib.query("wf_tb4_elems", 4 * L, d.window, d.num_nodes, self.G * self.Bw)), **f32)
            self.c = torch.empty(Ls, int(_lib.query("wf_tb4_elems", L, d.window, d.num_nodes, self.G * self.Bw)), **f32)
            # hidden states of every layer as fp16 hi/lo planes in the TB8 layout: the next layer's projection operand and
            # both weight-gradient operands.  Zero-initialised: the padding rows of a node tile are never written and must
            # read as zero where rows are contracted.  The head reads the top layer's last step from ``hlast`` (fp32).
            # h16: fp16, what the next layer's projection reads (masked under dropout); hb16: bf16, the plain h for the weight
            # gradients (one operand format per tensor-core product, and dG needs bf16's range); hb16m: bf16, masked
            self.hp = int(_lib.query("wf_tb8_elems", L, d.window, d.num_nodes, self.G * self.Bw))
            self.h16 = torch.zeros(max(Ls - 1, 1), 2, self.hp, **i16)
            self.hb16 = torch.zeros(Ls, 2, self.hp, **i16) if self.training else None
            self.hb16m = (torch.zeros(max(Ls - 1, 1), 2, self.hp, **i16)
                          if self.training and self.dropout[1] > 0 and Ls > 1 else None)
            self.hlast = torch.empty(self.W * d.num_nodes, L, **f32)
            self.feats16 = None   # layer-0 input planes when the features arrive as an fp32 tensor (module API)
            self.h = None
        else:
            self.gates = torch.empty(Ls, self.rows, 4 * L, **f32)
            self.c = torch.empty(Ls, self.rows, L, **f32)
            self.h = torch.empty(Ls, self.rows, L, **f32)
        # dropout state: {seed, forward-pass counter} on the device (kernels read it, wf_rng_advance bumps it on the stream)
        self.rng = torch.tensor([int(seed), 0], dtype=torch.int64, device=self.device)
        p_lstm, p_head = self.dropout[1], self.dropout[2]
        self.hlast_m = torch.empty(self.W * d.num_nodes, L, **f32) if p_head > 0 else None
        self.h_masked = None
        if p_lstm > 0 and not self.tc:
            self.h_masked = torch.empty(Ls - 1, self.rows, L, **f32)
        self.pred = torch.empty(self.W * d.num_nodes, d.O, **f32)
        self.dpred = torch.empty(self.W * d.num_nodes, d.O, **f32)
        self.loss = torch.zeros(self.W, **f32)
        self.dlast = torch.empty(self.W * d.num_nodes, L, **f32)
        self.P = d.P
        self.grads = torch.zeros(self.G, self.P, **f32)
        self.norms = torch.zeros(self.G, **f32)
        ws = max(_lib.query("wf_lstm_bwd_workspace_bytes", Ls, d.hidden, L, d.window, d.num_nodes, self.G, self.Bw),
                 _lib.query("wf_head_workspace_bytes", L, d.O, d.num_nodes, self.G, self.Bw),
                 _lib.query("wf_optim_workspace_bytes", self.G))
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.seq:
            # persistent path: 16-bit hi/lo operands everywhere (fp16 forward, bf16 for gradient operands)
            ws = max(ws, _lib.query("wf_lstm_bwd_seq_workspace_bytes", Ls, d.hidden, L, d.window, d.num_nodes, self.G,
                                    self.Bw))
            nw = int(_lib.query("wf_seq_weight_elems", Ls, L, self.G))
            self.PT = int(_lib.query("wf_param_count_transposed", Ls, d.hidden, L, d.O))
            self.w16 = [torch.empty(nw, **i16) for _ in range(4)]            # W_hh: fwd hi/lo, bwd hi/lo
            s16 = [int(_lib.query("wf_param_stride16", Ls, d.hidden, L, d.O, w)) for w in (0, 1)]
            self.p16 = [torch.empty(self.G, s16[0], **i16) for _ in range(2)]   # flat params as fp16 hi/lo
            self.pT16 = [torch.empty(self.G, s16[1], **i16) for _ in range(2)]  # W_ih^T as bf16 hi/lo
            # one layer's dG as bf16 hi/lo planes (TB8), zero-initialised like h16
            self.dg16 = (torch.zeros(2, int(_lib.query("wf_tb8_elems", 4 * L, d.window, d.num_nodes, self.W)), **i16)
                         if self.training else None)
            self._gcn_lo = {}
        elif self.tc:
            ws = max(ws, _lib.query("wf_lstm_bwd_tc_workspace_bytes", Ls, d.hidden, L, d.window, d.num_nodes, self.G,
                                    self.Bw))
            self.PT = int(_lib.query("wf_param_count_transposed", Ls, d.hidden, L, d.O))
            self.params_lo = torch.empty(self.G, self.P, **f32)
            self.paramsT = torch.empty(self.G, self.PT, **f32)
            self.paramsT_lo = torch.empty(self.G, self.PT, **f32)
            if self.training:
                rt = int(_lib.query("wf_transposed_pitch", d.window, d.num_nodes))
                self.hT = torch.zeros(Ls, self.W * L * rt, **f32)
                self.hT_lo = torch.zeros(Ls, self.W * L * rt, **f32)
                self.featsT = torch.zeros(self.W * d.hidden * rt, **f32)
                self.featsT_lo = torch.zeros(self.W * d.hidden * rt, **f32)
                self.dgT = torch.zeros(self.W * 4 * L * rt, **f32)
            else:
                self.hT = self.hT_lo = self.featsT = self.featsT_lo = self.dgT = None
            self._gcn_lo = {}
        self.ws_bytes = int(ws)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        # side stream for work that is off the critical path of a pass (weight staging next to the frozen GCN stack, the
        # head's parameter gradients next to the LSTM backward); WF_OVERLAP=0 keeps everything on one stream
        self.overlap = os.environ.get("WF_OVERLAP", "1") != "0"
        self._side = torch.cuda.Stream(self.device)
        self.ws_head = torch.empty(int(_lib.query("wf_head_workspace_bytes", L, d.O, d.num_nodes, self.G, self.Bw)),
                                   dtype=torch.uint8, device=self.device)
        self.feats = None
        self.agg = None  # scratch of the GCN pre-aggregation pass (persistent path)
        self.launches = 0  # kernels enqueued by this engine (bench.py reports it)

    def _join16(self, planes, fmt=0):
        """fp32 view (hi + lo) of a pair of 16-bit planes [2, n]."""
        n = planes[0].numel()
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("wf_join16", _lib.ptr(planes[0]), _lib.ptr(planes[1]), n, fmt, _lib.ptr(out), _lib.stream_ptr())
        return out

    def _tb8_to_rows(self, flat, channels):
        """TB8 [window, step, node tile, channels/8, 128 rows, 8] -> row-major [G*Bw*R, channels]."""
        d = self.dims
        tpw = (d.num_nodes + 127) // 128
        rpt = int(_lib.query("wf_tile_rows", d.num_nodes))  # nodes per tile; rows beyond are padding
        t = flat.view(self.W, d.window, tpw, channels // 8, 128, 8).permute(0, 1, 2, 4, 3, 5)
        t = t.reshape(self.W, d.window, tpw, 128, channels)[:, :, :, :rpt]
        return t.reshape(self.W, d.window, tpw * rpt, channels)[:, :, :d.num_nodes].reshape(self.rows, channels)

    def hidden_states(self, masked=False):
        """[layers, G*Bw*R, L] row-major fp32 copy of the LSTM hidden states, whatever layout the kernels keep them in
        (``masked``: what the next layer read when inter-layer dropout is on; layers - 1 entries)."""
        d, Ls, L = self.dims, self.dims.lstm_layers, self.dims.lstm_hidden
        if not self.seq:
            return (self.h_masked if masked else self.h).clone()
        if masked or self.hb16 is None:  # fp16 planes of what the next layer read (layers - 1 entries)
            return torch.stack([self._tb8_to_rows(self._join16(self.h16[l]), L) for l in range(Ls - 1)])
        return torch.stack([self._tb8_to_rows(self._join16(self.hb16[l], 1), L) for l in range(Ls)])

    def gcn_features(self, layer=None):
        """fp32 [G*Bw*R, hidden] copy of the GCN stack's output (``layer`` 0..3 with keep_gcn_activations)."""
        if not self.gcn_ss:
            return (self.feats if layer is None else self.act[layer]).clone()
        src = self.feats if layer is None else self.act16[layer]
        return self._join16(src).view(self.rows, self.dims.hidden)

    def train(self, mode=True):
        """Dropout on (if any p > 0) / off, like ``nn.Module.train()``."""
        self.stochastic = bool(mode) and any(p > 0 for p in self.dropout)
        return self

    def eval(self):
        return self.train(False)

    def _p(self, site):
        return self.dropout[site] if self.stochastic else 0.0

    @_on_device
    def advance_rng(self):
        """Next forward pass draws fresh masks (enqueued on the stream: also advances under CUDA-graph replay)."""
        _lib.call("wf_rng_advance", _lib.ptr(self.rng), _lib.stream_ptr())
        self.launches += 1

    def check(self):
        """Synchronise and raise if a kernel flagged an error: a tensor-core pipeline wait that timed out (never
        expected) or an activation outside the fp16 operand range."""
        raise_on_error_code(int(self.err.item()))

    @_on_device
    def _gcn_w_lo(self, Wt):
        """Cached operand staging of a GCN weight: TF32 lo half (stepwise path) or fp16 (hi, lo) (persistent path)."""
        # the entry keeps the weight tensor alive: an address can then not be reused by another tensor while it is cached
        key = (Wt.data_ptr(), Wt._version)
        hit = self._gcn_lo.get(key)
        lo = hit[1] if hit is not None and hit[0] is Wt else None
        if lo is None:
            if self.seq:
                lo = (torch.empty(Wt.shape, dtype=torch.int16, device=Wt.device),
                      torch.empty(Wt.shape, dtype=torch.int16, device=Wt.device))
                _lib.call("wf_split16", _lib.ptr(Wt), _lib.ptr(lo[0]), _lib.ptr(lo[1]), Wt.numel(), 0, _lib.stream_ptr())
            else:
                lo = torch.empty_like(Wt)
                _lib.call("wf_split_lo", _lib.ptr(Wt), _lib.ptr(lo), Wt.numel(), _lib.stream_ptr())
            if len(self._gcn_lo) > 16:
                self._gcn_lo.clear()
            self._gcn_lo[key] = (Wt, lo)
        return lo

    # ------------------------------------------------------------------ GCN stack
    @_on_device
    def gcn_forward(self, X, x_ld, x_win_stride, x_win_off, gcn_weights, graphs):
        """4 x relu(GCNConv), dropout after all layers but the last (hybrid_model.py:65-76); returns the
        [G*Bw*R, hidden] feature buffer."""
        d, st = self.dims, _lib.stream_ptr()
        if isinstance(graphs, RegionGraph):
            rp, cl, vl, rps, cs = graphs.rowptr, graphs.col, graphs.val, 0, 0
            gl, gmax, gls = graphs.gather_rows, int(graphs.gather_rows.numel()), 0
        elif isinstance(graphs, StackedGraphs):
            rp, cl, vl, rps, cs = graphs.rowptr, graphs.col, graphs.val, graphs.rowptr_stride, graphs.csr_stride
            gl, gmax, gls = graphs.gather_rows, graphs.gather_max, graphs.gather_rows.shape[1]
        else:
            raise TypeError("graphs must be a RegionGraph or StackedGraphs")
        if graphs.R != d.R:
            raise ValueError(f"graph was normalised over {graphs.R} rows, engine window has {d.R}")
        src, src_ld, src_stride, src_off, cin = X, x_ld, x_win_stride, x_win_off, d.in_channels
        nlayers = len(gcn_weights)
        if self.gcn_ss and x_ld == cin and (x_win_off is not None or x_win_stride == d.R * cin):
            return self._gcn_forward_ss(X, x_win_off, gcn_weights, graphs, rp, cl, vl, rps, cs)
        if self.gcn_ss:
            raise ValueError("the persistent path reads windows as contiguous [R, in_channels] slices (dataset.py:36-37)")
        for i, (Wt, b) in enumerate(gcn_weights):
            dst = self.act[i] if self.keep_gcn else self.act[i & 1]
            dense = src_off is None and src_ld == cin and src_stride == d.R * cin
            p_drop = self._p(0) if i < nlayers - 1 else 0.0
            if self.tc and not self.seq and dense and cin % 32 == 0:
                want_t = self.training and i == nlayers - 1
                _lib.call("wf_gcn_layer_fwd_tc", _lib.ptr(src), _lib.ptr(Wt), _lib.ptr(self._gcn_w_lo(Wt)), _lib.ptr(b),
                          _lib.ptr(rp), _lib.ptr(cl), _lib.ptr(vl), rps, cs, d.R, d.num_nodes, cin, d.hidden, self.G,
                          self.Bw, 1, _lib.ptr(dst), _lib.ptr(self.featsT) if want_t else None,
                          _lib.ptr(self.featsT_lo) if want_t else None, _lib.ptr(self.err), st)
            else:
                _lib.call("wf_gcn_layer_fwd", _lib.ptr(src), src_ld, src_stride, _lib.ptr(src_off), _lib.ptr(Wt),
                          _lib.ptr(b), 0, 0, _lib.ptr(rp), _lib.ptr(cl), _lib.ptr(vl), rps, cs, d.R, cin, d.hidden,
                          self.G, self.Bw, 1, _lib.ptr(dst), st)
            self.launches += 1
            if p_drop > 0:  # paths without a fused epilogue mask: the same mask as a separate pass
                _lib.call("wf_dropout_apply", _lib.ptr(dst), 0, self.rows, d.hidden, self.rows, d.hidden, p_drop,
                          _lib.ptr(self.rng), i, _lib.ptr(dst), st)
                self.launches += 1
            src, src_ld, src_stride, src_off, cin = dst, d.hidden, d.R * d.hidden, None, d.hidden
        self.feats = src
        return src

    def _gcn_forward_ss(self, X, x_win_off, gcn_weights, graphs, rp, cl, vl, rps, cs):
        """The stack on pre-split fp16 hi/lo planes: per layer [split of the fp32 windows (first layer only)] + aggregation
        of the leading rows that have neighbours + one persistent SS-mode GEMM with resident weights whose epilogue writes
        the next layer's operand planes through TMA stores (csrc/wf_gemm_ss.cu)."""
        d, st = self.dims, _lib.stream_ptr()
        agg_rows = int(graphs.agg_rows)
        if agg_rows > 0 and (self.side16 is None or self.side16.numel() < 2 * self.W * agg_rows * d.hidden):
            self.side16 = torch.empty(2 * self.W * agg_rows * d.hidden, dtype=torch.int16, device=self.device)
        nlayers, cin, src16 = len(gcn_weights), d.in_channels, None
        for i, (Wt, b) in enumerate(gcn_weights):
            dst = self.act16[i] if self.keep_gcn else self.act16[i & 1]
            p_drop = self._p(0) if i < nlayers - 1 else 0.0
            w16 = self._gcn_w_lo(Wt)
            _lib.call("wf_gcn_layer_fwd_ss", _lib.ptr(X) if i == 0 else None, _lib.ptr(x_win_off) if i == 0 else None,
                      _lib.ptr(src16) if i > 0 else None, _lib.ptr(self.x16) if i == 0 else None, _lib.ptr(w16[0]),
                      _lib.ptr(w16[1]), _lib.ptr(b), _lib.ptr(rp), _lib.ptr(cl), _lib.ptr(vl), rps, cs, agg_rows,
                      _lib.ptr(self.side16) if agg_rows > 0 else None, d.R, cin, d.hidden, self.G, self.Bw, 1,
                      _lib.ptr(dst), _lib.ptr(self.featsb16) if i == nlayers - 1 else None, p_drop, _lib.ptr(self.rng), i,
                      _lib.ptr(self.err), st)
            self.launches += 2 + (1 if i == 0 else 0) if agg_rows > 0 else 1 + (1 if i == 0 else 0)
            src16, cin = dst, d.hidden
        self.feats = src16
        self._xb16 = self.featsb16
        return src16

    # ------------------------------------------------------------------ LSTM + head
    def _seq_input_planes(self, feats):
        """Layer-0 operand of the persistent path: the GCN stack's fp16 hi/lo planes, or -- for features that arrive as
        an fp32 tensor (drop-in module API) -- their split into ``feats16``."""
        d = self.dims
        if feats.dtype == torch.int16:
            return feats
        if self.feats16 is None:
            self.feats16 = torch.empty(2, 2, self.rows, d.hidden, dtype=torch.int16, device=self.device)
        for fmt in ((0, 1) if self.training else (0,)):  # fp16 for the projection, bf16 for the weight gradient
            _lib.call("wf_split16", _lib.ptr(feats), _lib.ptr(self.feats16[fmt, 0]), _lib.ptr(self.feats16[fmt, 1]),
                      feats.numel(), fmt, _lib.stream_ptr())
            self.launches += 1
        self._xb16 = self.feats16[1]
        return self.feats16[0]

    @_on_device
    def _prep_weights(self, params, params_stride):
        """Operand staging of the persistent path: 16-bit hi/lo copies of the current (fast) weights, on the current stream."""
        d = self.dims
        src_stride = params_stride if self.G > 1 else self.P
        _lib.call("wf_prep_weights_seq", _lib.ptr(params), src_stride, d.lstm_layers, d.hidden, d.lstm_hidden, d.O, self.G,
                  _lib.ptr(self.p16[0]), _lib.ptr(self.p16[1]), _lib.ptr(self.pT16[0]), _lib.ptr(self.pT16[1]),
                  _lib.ptr(self.w16[0]), _lib.ptr(self.w16[1]), _lib.ptr(self.w16[2]), _lib.ptr(self.w16[3]), _lib.stream_ptr())

    @_on_device
    def lstm_head_forward(self, params, params_stride, feats=None, prepped=False):
        d, st = self.dims, _lib.stream_ptr()
        feats = self.feats if feats is None else feats
        Ls, L = d.lstm_layers, d.lstm_hidden
        p_lstm, p_head = self._p(1), self._p(2)
        if self.seq:
            if not prepped:
                self._prep_weights(params, params_stride)
            self._x16 = self._seq_input_planes(feats)
            _lib.call("wf_lstm_fwd_seq", _lib.ptr(self._x16), _lib.ptr(params), _lib.ptr(self.p16[0]), _lib.ptr(self.p16[1]),
                      params_stride if self.G > 1 else self.P, _lib.ptr(self.w16[0]), _lib.ptr(self.w16[1]), Ls, d.hidden,
                      L, d.O, d.window, d.num_nodes, self.G, self.Bw, _lib.ptr(self.gates), _lib.ptr(self.h16),
                      _lib.ptr(self.c), _lib.ptr(self.hlast), _lib.ptr(self.hb16), p_lstm, _lib.ptr(self.rng),
                      _lib.ptr(self.hb16m), _lib.ptr(self.err), st)
            self.launches += (1 + (Ls - 1) + 1) + 2 * Ls  # operand staging, then (projection + recurrence) per layer
            h_top, t_head = self.hlast, 1   # the top layer's last step, compact: the head sees windows of ONE step
        elif self.tc:
            # operand staging for 3xTF32: lo halves and transposed copies of the current weights
            src_stride = params_stride if self.G > 1 else self.P
            _lib.call("wf_prep_weights_tc", _lib.ptr(params), src_stride, Ls, d.hidden, L, d.O, self.G,
                      _lib.ptr(self.params_lo), _lib.ptr(self.paramsT), _lib.ptr(self.paramsT_lo), st)
            _lib.call("wf_lstm_fwd_tc", _lib.ptr(feats), _lib.ptr(params), _lib.ptr(self.params_lo), params_stride, Ls,
                      d.hidden, L, d.O, d.window, d.num_nodes, self.G, self.Bw, _lib.ptr(self.gates), _lib.ptr(self.h),
                      _lib.ptr(self.c), _lib.ptr(self.hT), _lib.ptr(self.hT_lo), _lib.ptr(self.err), st)
            self.launches += 2 * Ls + Ls * (1 + d.window)
            h_top, t_head = self.h[Ls - 1], d.window
        else:
            _lib.call("wf_lstm_fwd", _lib.ptr(feats), _lib.ptr(params), params_stride, Ls, d.hidden, L, d.O, d.window,
                      d.num_nodes, self.G, self.Bw, _lib.ptr(self.gates), _lib.ptr(self.h), _lib.ptr(self.c), p_lstm,
                      _lib.ptr(self.rng), _lib.ptr(self.h_masked), st)
            self.launches += Ls * (1 + d.window) + (Ls - 1 if p_lstm > 0 else 0)
            h_top, t_head = self.h[Ls - 1], d.window
        if p_head > 0:
            # head-input dropout (hybrid_model.py:108): mask the last step of the top layer into a compact [W*N, L]
            # buffer, which the head then reads as a window of ONE step
            last = h_top.view(-1)[(t_head - 1) * d.num_nodes * L:]
            _lib.call("wf_dropout_apply", _lib.ptr(last), t_head * d.num_nodes * L, d.num_nodes, L, self.W * d.num_nodes, L,
                      p_head, _lib.ptr(self.rng), SITE_HEAD, _lib.ptr(self.hlast_m), st)
            h_top, t_head = self.hlast_m, 1
            self.launches += 1
        self._h_top, self._t_head = h_top, t_head
        _lib.call("wf_head_fwd", _lib.ptr(h_top), _lib.ptr(params), params_stride, Ls, d.hidden, L, d.O, t_head,
                  d.num_nodes, self.G, self.Bw, _lib.ptr(self.pred), st)
        self.launches += 1
        return self.pred

    @_on_device
    def mse(self, y=None, feat=None, tgt_off=None, feat_ld=0, grad_scale=1.0, want_grad=True):
        """Per-window nn.MSELoss (+ backward seed) against explicit y or in-place targets."""
        d = self.dims
        _lib.call("wf_mse_fwd_bwd", _lib.ptr(self.pred), _lib.ptr(y), _lib.ptr(feat), _lib.ptr(tgt_off), feat_ld,
                  d.num_weather, d.num_nodes, d.O, self.W, float(grad_scale), _lib.ptr(self.loss),
                  _lib.ptr(self.dpred) if want_grad else None, _lib.stream_ptr())
        self.launches += 1
        return self.loss

    @_on_device
    def backward(self, params, params_stride, feats=None, dpred=None):
        """BPTT from ``dpred`` (default: the seed left by ``mse``) into ``self.grads`` [G, P].  With dropout on, the
        masks of the forward pass are regenerated from (seed, pass counter): call before ``advance_rng()``."""
        d, st = self.dims, _lib.stream_ptr()
        feats = self.feats if feats is None else feats
        dpred = self.dpred if dpred is None else dpred
        Ls, L = d.lstm_layers, d.lstm_hidden
        p_lstm, p_head = self._p(1), self._p(2)
        # the head's input is whatever the forward pass fed it: the (masked) compact last step or the top layer's rows
        forked = self.seq and self.overlap
        if forked:
            # dW_o / db_o are needed only by the optimiser step: on the side stream, next to the LSTM backward, whose
            # recurrence kernels leave 28 SMs idle; the BPTT chain only waits for dlast
            main = torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                _lib.call("wf_head_bwd", _lib.ptr(dpred), _lib.ptr(self._h_top), _lib.ptr(params), params_stride, Ls,
                          d.hidden, L, d.O, self._t_head, d.num_nodes, self.G, self.Bw, None, _lib.ptr(self.grads), self.P,
                          _lib.ptr(self.ws_head), self.ws_head.numel(), _lib.stream_ptr())
        _lib.call("wf_head_bwd", _lib.ptr(dpred), _lib.ptr(self._h_top), _lib.ptr(params), params_stride, Ls,
                  d.hidden, L, d.O, self._t_head, d.num_nodes, self.G, self.Bw, _lib.ptr(self.dlast),
                  None if forked else _lib.ptr(self.grads), self.P, _lib.ptr(self.ws), self.ws_bytes, st)
        if p_head > 0:
            _lib.call("wf_dropout_apply", _lib.ptr(self.dlast), 0, self.W * d.num_nodes, L, self.W * d.num_nodes, L, p_head,
                      _lib.ptr(self.rng), SITE_HEAD, _lib.ptr(self.dlast), st)
            self.launches += 1
        if self.tc:
            if not self.training:
                raise RuntimeError("engine was built with training=False")
            if self.seq:
                _lib.call("wf_lstm_bwd_seq", _lib.ptr(self._xb16), _lib.ptr(self.pT16[0]), _lib.ptr(self.pT16[1]),
                          _lib.ptr(self.w16[2]), _lib.ptr(self.w16[3]), Ls, d.hidden, L, d.O, d.window, d.num_nodes, self.G,
                          self.Bw, _lib.ptr(self.gates), _lib.ptr(self.c), _lib.ptr(self.hb16), _lib.ptr(self.dg16),
                          _lib.ptr(self.dlast), _lib.ptr(self.grads), self.P, p_lstm, _lib.ptr(self.rng),
                          _lib.ptr(self.hb16m), _lib.ptr(self.ws), self.ws_bytes, _lib.ptr(self.err), st)
                self.launches += 6 + Ls * 4 + 1  # head bwd + per layer: recurrence, wgrad + reduce, dX (layer 0: two wgrads)
            else:
                _lib.call("wf_lstm_bwd_tc", _lib.ptr(self.featsT), _lib.ptr(self.featsT_lo), _lib.ptr(self.paramsT),
                          _lib.ptr(self.paramsT_lo), Ls, d.hidden, L, d.O, d.window, d.num_nodes, self.G, self.Bw,
                          _lib.ptr(self.gates), _lib.ptr(self.c), _lib.ptr(self.hT), _lib.ptr(self.hT_lo),
                          _lib.ptr(self.dgT), _lib.ptr(self.dlast), _lib.ptr(self.grads), self.P, _lib.ptr(self.ws),
                          self.ws_bytes, _lib.ptr(self.err), st)
                self.launches += 6 + Ls * (d.window + 5)
        else:
            _lib.call("wf_lstm_bwd", _lib.ptr(feats), _lib.ptr(params), params_stride, Ls, d.hidden, L, d.O, d.window,
                      d.num_nodes, self.G, self.Bw, _lib.ptr(self.gates), _lib.ptr(self.h), _lib.ptr(self.c),
                      _lib.ptr(self.dlast), _lib.ptr(self.grads), self.P, p_lstm, _lib.ptr(self.rng),
                      _lib.ptr(self.h_masked), _lib.ptr(self.ws), self.ws_bytes, st)
            self.launches += 6 + Ls * (d.window + 9) + (Ls - 1 if p_lstm > 0 else 0)
        if forked:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        return self.grads

    @_on_device
    def sgd_step(self, fast, lr, max_norm=1.0):
        """fast[g] -= lr * clip(grads[g]) for every task (train_hybrid_maml_v5.py:135-139)."""
        _lib.call("wf_clip_sgd_step", _lib.ptr(fast), self.P, _lib.ptr(self.grads), self.P, self.P, self.G, float(lr),
                  float(max_norm), _lib.ptr(self.norms), _lib.ptr(self.ws), self.ws_bytes, _lib.stream_ptr())
        self.launches += 2
        return fast

    # ------------------------------------------------------------------ convenience
    def forward_backward(self, X, x_ld, x_win_stride, x_win_off, gcn_weights, graphs, params, params_stride,
                         y=None, feat=None, tgt_off=None, feat_ld=0, grad_scale=1.0):
        if self.seq and self.overlap:
            # the operand staging of the (fast) LSTM weights does not depend on the frozen GCN stack: fork it onto the side
            # stream (also inside a captured graph: the fork / join become graph edges)
            main = torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self._prep_weights(params, params_stride)
            self.gcn_forward(X, x_ld, x_win_stride, x_win_off, gcn_weights, graphs)
            main.wait_stream(self._side)
            self.lstm_head_forward(params, params_stride, prepped=True)
        else:
            self.gcn_forward(X, x_ld, x_win_stride, x_win_off, gcn_weights, graphs)
            self.lstm_head_forward(params, params_stride)
        self.mse(y, feat, tgt_off, feat_ld, grad_scale)
        self.backward(params, params_stride)
        if self.stochastic:
            self.advance_rng()
        return self.loss, self.grads


SITE_HEAD = 32  # csrc/wf_rng.cuh: GCN layer i = i, LSTM layer l = 16 + l, head input = 32
SITE_LSTM = 16


def raise_on_error_code(code):
    if code == 0:
        return
    if code in ERR_MESSAGES:
        raise RuntimeError(f"wf_stgcn device error {code}: {ERR_MESSAGES[code]}")
    raise RuntimeError(f"tcgen05 pipeline timeout (role code {code})")


class AdamState:
    """Flat Adam/AdamW state + the 8-float hyper-parameter block the kernel reads from HBM."""

    def __init__(self, P, device, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True):
        self.P, self.device = int(P), torch.device(device)
        self.exp_avg = torch.zeros(P, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(P, dtype=torch.float32, device=device)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.decoupled, self.step_count = bool(decoupled), 0
        # ring of pinned staging slots: the async H2D of step t must not be overwritten by the
        # host preparing step t+1 while the stream still lags behind
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.hyper_host = torch.zeros(2048, 8, dtype=torch.float32).pin_memory()
        self.hyper = torch.zeros(8, dtype=torch.float32, device=device)
        self.norm = torch.zeros(1, dtype=torch.float32, device=device)
        self.ws_bytes = int(_lib.query("wf_optim_workspace_bytes", 1))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)

    @_on_device
    def prepare(self, grad_scale=1.0):
        """Host half of a step: count it and send {lr, betas, eps, weight decay, bias corrections, grad scale} to the
        device block the kernel reads (async copy from a pinned slot, stream ordered).  Separate from ``apply`` so that
        the kernel launch can live inside a captured CUDA graph while the numbers still change every step."""
        self.step_count += 1
        b1, b2 = self.betas
        slot = self.hyper_host[self.step_count % self.hyper_host.shape[0]]
        slot.copy_(torch.tensor([self.lr, b1, b2, self.eps, self.weight_decay, 1.0 - b1 ** self.step_count,
                                 1.0 - b2 ** self.step_count, grad_scale], dtype=torch.float32))
        self.hyper.copy_(slot, non_blocking=True)

    @_on_device
    def apply(self, theta, grad, max_norm=1.0):
        """Device half: fused clip + Adam(W) with the hyper-parameters currently in ``self.hyper`` (capturable)."""
        _lib.call("wf_clip_adam_step", _lib.ptr(theta), _lib.ptr(grad), _lib.ptr(self.exp_avg),
                  _lib.ptr(self.exp_avg_sq), self.P, _lib.ptr(self.hyper), float(max_norm), int(self.decoupled),
                  _lib.ptr(self.norm), _lib.ptr(self.ws), self.ws_bytes, _lib.stream_ptr())
        return theta

    def step(self, theta, grad, max_norm=1.0, grad_scale=1.0):
        self.prepare(grad_scale)
        return self.apply(theta, grad, max_norm)
