"""torch.autograd.Functions over the C-ABI launchers (single parameter set, G = 1).

These back the drop-in ``nn.Module`` classes (model.py, hybrid_model.py) so that user code
written against the reference -- ``out = model(batch.x, batch.edge_index); loss.backward()`` --
runs unchanged on the CUDA kernels.  The task-batched MAML path does not go through autograd
at all (engine.py / train_hybrid_maml_v5.py).

Kernels.  With ``precision() == "tf32x3"`` (the default) the modules run on the same tcgen05
kernels as the task-batched engine whenever the shapes allow it: ``GCNConvReLU`` on the
persistent fp16 hi/lo GEMM (``wf_gcn_layer_fwd_g16``: Cout % 128 == 0, Cin % 8 == 0) and
``LSTMHead`` on the persistent recurrence kernels (``wf_lstm_fwd_seq`` / ``wf_lstm_bwd_seq``:
hidden 128, input width % 128 == 0) through a leased ``HybridEngine``; other shapes, and
``set_precision("fp32")``, take the exact-FP32 CUDA-core kernels.  There is no PyTorch fallback.

Dropout (train mode, p > 0) is fused: counter-based masks (csrc/wf_rng.cuh) drawn in the GCN
epilogue / recurrence kernel / head pass and regenerated in backward from the (seed, pass)
pair snapshotted at forward time.  The seed comes from torch's default generator when a device
is first used, so ``torch.manual_seed`` makes runs repeatable.
"""
from __future__ import annotations

import torch

from . import _lib
from .engine import HybridEngine, raise_on_error_code
from .graph import RegionGraph

_PRECISION = "tf32x3"


def set_precision(mode):
    """"tf32x3": tensor-core kernels where the shapes allow (default); "fp32": exact-FP32 CUDA-core kernels."""
    global _PRECISION
    if mode not in ("tf32x3", "fp32"):
        raise ValueError("precision must be 'tf32x3' or 'fp32'")
    _PRECISION = mode


def precision():
    return _PRECISION


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------ per-device state
class _DeviceState:
    """RNG pass counter, error flag (with a pinned mirror read without synchronising) and operand caches."""

    def __init__(self, device):
        self.device = device
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # torch.manual_seed -> repeatable masks
        self.rng = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.engines = {}

    def rng_snapshot(self):
        """(seed, pass) of THIS call as its own device tensor; the shared counter moves on (all on the stream)."""
        snap = self.rng.clone()
        _lib.call("wf_rng_advance", _lib.ptr(self.rng), _lib.stream_ptr())
        return snap

    def poll_errors(self):
        """Raise if a kernel of an EARLIER call flagged an error (the mirror is filled asynchronously)."""
        code = int(self.err_host[0])
        if code != 0:
            self.err_host.zero_()
            self.err.zero_()
            raise_on_error_code(code)

    def publish_errors(self):
        self.err_host.copy_(self.err, non_blocking=True)



def split_weight16(w):
    """fp16 (hi, lo) operand halves of a weight matrix (wf_split16).  Callers that own the weight cache the pair
    themselves, keyed by (data_ptr, _version) -- see model.GCNConv; a device-wide cache keyed by address would hand a
    new tensor allocated at a freed address somebody else's operands."""
    with torch.cuda.device(w.device):
        hi = torch.empty(w.shape, dtype=torch.int16, device=w.device)
        lo = torch.empty(w.shape, dtype=torch.int16, device=w.device)
        _lib.call("wf_split16", _lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), w.numel(), 0, _lib.stream_ptr())
    return hi, lo


_STATES = {}


def _state(device):
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    st = _STATES.get(device)
    if st is None:
        with torch.cuda.device(device):
            st = _STATES[device] = _DeviceState(device)
    return st


def check(device=None):
    """Synchronise and raise if any kernel launched through this module flagged an error."""
    for dev, st in list(_STATES.items()):
        if device is None or torch.device(device) == dev:
            code = int(st.err.item())
            if code:
                st.err.zero_()
                raise_on_error_code(code)


# ------------------------------------------------------------------------------------------ GCN layer
class GCNConvReLU(torch.autograd.Function):
    """Y = dropout([relu]((A_hat X) W^T + b)) for one or more windows sharing a graph.

    forward: wf_gcn_layer_fwd_g16 (tensor cores) or wf_gcn_layer_fwd (FP32); backward: wf_gcn_layer_bwd_ss (tensor
    cores: dZ written once as bf16 hi/lo planes, dW / db by the MN-major weight-gradient kernel, dX by the SS GEMM +
    transposed aggregation) or wf_gcn_layer_bwd (FP32, and widths the tensor-core path does not cover).
    ``p_drop`` > 0 applies nn.Dropout after the ReLU (model.py:33-42) as site ``site`` (the layer index)."""

    @staticmethod
    def forward(ctx, x, weight, bias, graph: RegionGraph, relu: bool, p_drop: float = 0.0, site: int = 0, w16=None):
        _lib.require_cuda(x, weight, bias)
        x, weight, bias = _f32c(x), _f32c(weight), _f32c(bias)
        rows, cin = x.shape
        cout = weight.shape[0]
        if rows % graph.R != 0:
            raise ValueError(f"x has {rows} rows, graph was normalised over {graph.R}")
        bw = rows // graph.R
        y = torch.empty(rows, cout, dtype=torch.float32, device=x.device)
        p_drop = float(p_drop)
        with torch.cuda.device(x.device):
            st = _state(x.device)
            st.poll_errors()
            snap = st.rng_snapshot() if p_drop > 0 else None
            s = _lib.stream_ptr()
            fast = _PRECISION == "tf32x3" and cout % 128 == 0 and cin % 8 == 0 and x.data_ptr() % 16 == 0
            if fast:
                hi, lo = w16 if w16 is not None else split_weight16(weight.detach())
                gl = graph.gather_rows
                gmax = int(gl.numel())
                agg = torch.empty(rows, cin, dtype=torch.float32, device=x.device) if gmax > 0 else None
                _lib.call("wf_gcn_layer_fwd_g16", _lib.ptr(x), None, rows, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(bias),
                          _lib.ptr(graph.rowptr), _lib.ptr(graph.col), _lib.ptr(graph.val), 0, 0,
                          _lib.ptr(gl) if gmax > 0 else None, gmax, 0, _lib.ptr(agg), graph.R, graph.R, cin, cout, 1, bw,
                          int(relu), _lib.ptr(y), None, None, p_drop, _lib.ptr(snap), int(site), _lib.ptr(st.err), s)
                st.publish_errors()
            else:
                _lib.call("wf_gcn_layer_fwd", _lib.ptr(x), cin, graph.R * cin, None, _lib.ptr(weight), _lib.ptr(bias), 0, 0,
                          _lib.ptr(graph.rowptr), _lib.ptr(graph.col), _lib.ptr(graph.val), 0, 0, graph.R, cin, cout, 1,
                          bw, int(relu), _lib.ptr(y), s)
                if p_drop > 0:
                    _lib.call("wf_dropout_apply", _lib.ptr(y), 0, rows, cout, rows, cout, p_drop, _lib.ptr(snap), int(site),
                              _lib.ptr(y), s)
        ctx.graph, ctx.relu, ctx.bw, ctx.p_drop, ctx.site, ctx.snap = graph, bool(relu), bw, p_drop, int(site), snap
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        graph, bw = ctx.graph, ctx.bw
        rows, cin = x.shape
        cout = weight.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(weight) if need_w else None
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if need_b else None
        fast = (_PRECISION == "tf32x3" and cout % 128 == 0 and cin % 8 == 0 and cin <= 256 and (not need_x or cin % 128 == 0)
                and (need_w or not need_b))
        if fast:
            dy = _f32c(dy)
            nbytes = int(_lib.query("wf_gcn_layer_bwd_ss_workspace_bytes", graph.R, cin, cout, bw))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            with torch.cuda.device(x.device):
                st = _state(x.device)
                st.poll_errors()
                _lib.call("wf_gcn_layer_bwd_ss", _lib.ptr(x), _lib.ptr(y), _lib.ptr(dy), _lib.ptr(weight), _lib.ptr(graph.rowptr),
                          _lib.ptr(graph.col), _lib.ptr(graph.val), _lib.ptr(graph.rowptr_t), _lib.ptr(graph.col_t),
                          _lib.ptr(graph.val_t), graph.R, cin, cout, bw, int(ctx.relu), ctx.p_drop, _lib.ptr(ctx.snap), ctx.site,
                          _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), nbytes, _lib.ptr(st.err), _lib.stream_ptr())
                st.publish_errors()
            return dx, dw, db, None, None, None, None, None
        dy = _f32c(dy).clone()  # overwritten with dY * mask * (Y > 0)
        nbytes = int(_lib.query("wf_gcn_layer_bwd_workspace_bytes", graph.R, cin, cout, 1, bw))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            s = _lib.stream_ptr()
            if ctx.p_drop > 0:  # the forward mask again, from the (seed, pass) pair of that call
                _lib.call("wf_dropout_apply", _lib.ptr(dy), 0, rows, cout, rows, cout, ctx.p_drop, _lib.ptr(ctx.snap),
                          ctx.site, _lib.ptr(dy), s)
            _lib.call("wf_gcn_layer_bwd", _lib.ptr(x), cin, graph.R * cin, None, _lib.ptr(y), _lib.ptr(dy),
                      _lib.ptr(weight), 0, _lib.ptr(graph.rowptr), _lib.ptr(graph.col), _lib.ptr(graph.val),
                      _lib.ptr(graph.rowptr_t), _lib.ptr(graph.col_t), _lib.ptr(graph.val_t), 0, 0, graph.R, cin, cout,
                      1, bw, int(ctx.relu), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), 0, 0, _lib.ptr(ws), nbytes, s)
        return dx, dw, db, None, None, None, None, None


# ------------------------------------------------------------------------------------------ LSTM + head
class _Lease:
    """An engine checked out of the per-device pool between a forward and its backward."""

    def __init__(self, engine):
        self.engine, self.consumed = engine, False
        engine._wf_busy = True

    def release(self):
        if self.engine is not None:
            self.engine._wf_busy = False
            self.engine = None

    def __del__(self):
        self.release()


def _lease_engine(st, dims, bw, dropout, training):
    key = (dims, int(bw), _PRECISION, bool(training))
    pool = st.engines.setdefault(key, [])
    for e in pool:
        if not getattr(e, "_wf_busy", False) and e.dropout == (0.0, dropout[0], dropout[1]):
            return _Lease(e)
    e = HybridEngine(dims, 1, bw, st.device, precision=_PRECISION, training=training,
                     dropout=(0.0, dropout[0], dropout[1]))
    e.err = st.err  # one flag per device
    if len(pool) < 4:   # engines beyond the pool are simply garbage collected after use
        pool.append(e)
    return _Lease(e)


class LSTMHead(torch.autograd.Function):
    """feats [bw*T*N, F] -> predictions [bw*N, O]: multi-layer LSTM over every node + Linear head.

    ``flat`` is the trainable flat buffer (engine.trainable_layout order); its gradient comes
    back flat as well.  Features get no gradient: the reference detaches them
    (hybrid_model.py:63).  ``dropout`` = (p_lstm, p_head), applied when > 0."""

    @staticmethod
    def forward(ctx, feats, flat, dims, bw, dropout=(0.0, 0.0)):
        _lib.require_cuda(feats, flat)
        feats, flat = _f32c(feats), _f32c(flat)
        d = dims
        rows = bw * d.R
        if feats.shape != (rows, d.hidden):
            raise ValueError(f"features {tuple(feats.shape)} do not match [{rows}, {d.hidden}]")
        dropout = (float(dropout[0]), float(dropout[1]))
        need_grad = bool(ctx.needs_input_grad[1])  # False under torch.no_grad() and for detached weights
        with torch.cuda.device(feats.device):
            st = _state(feats.device)
            st.poll_errors()
            lease = _lease_engine(st, d, bw, dropout, training=True)
            e = lease.engine
            e.train(any(p > 0 for p in dropout))
            if e.stochastic:
                e.rng.copy_(st.rng_snapshot())
            pred = e.lstm_head_forward(flat, 0, feats=feats).clone()
            st.publish_errors()
        if need_grad:
            ctx.lease, ctx.dims, ctx.bw = lease, d, bw
            ctx.save_for_backward(feats, flat)
        else:
            lease.release()
        return pred

    @staticmethod
    def backward(ctx, dpred):
        feats, flat = ctx.saved_tensors
        lease = ctx.lease
        if lease.consumed or lease.engine is None:
            raise RuntimeError("LSTMHead.backward ran twice: BPTT overwrites the saved activations, call forward again")
        e = lease.engine
        with torch.cuda.device(feats.device):
            st = _state(feats.device)
            grads = e.backward(flat, 0, feats=feats, dpred=_f32c(dpred))[0].clone()
            st.publish_errors()
        lease.consumed = True
        lease.release()
        return None, grads, None, None, None


class LinearRows(torch.autograd.Function):
    """y = x W^T + b on the exact-FP32 GEMM kernels: the GCN layer launchers with identity aggregation
    (rowptr == NULL).  Backs ``STGCN.output_layer`` (model.py:49; 256 -> 96 on the last time slice only)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_cuda(x, weight, bias)
        x, weight, bias = _f32c(x), _f32c(weight), _f32c(bias)
        rows, cin = x.shape
        cout = weight.shape[0]
        y = torch.empty(rows, cout, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("wf_gcn_layer_fwd", _lib.ptr(x), cin, rows * cin, None, _lib.ptr(weight), _lib.ptr(bias), 0, 0,
                      None, None, None, 0, 0, rows, cin, cout, 1, 1, 0, _lib.ptr(y), _lib.stream_ptr())
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        rows, cin = x.shape
        cout = weight.shape[0]
        dy = _f32c(dy)
        need_x, need_w, need_b = ctx.needs_input_grad
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(weight) if need_w else None
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if need_b else None
        nbytes = int(_lib.query("wf_gcn_layer_bwd_workspace_bytes", rows, cin, cout, 1, 1))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("wf_gcn_layer_bwd", _lib.ptr(x), cin, rows * cin, None, None, _lib.ptr(dy), _lib.ptr(weight), 0,
                      None, None, None, None, None, None, 0, 0, rows, cin, cout, 1, 1, 0, _lib.ptr(dx), _lib.ptr(dw),
                      _lib.ptr(db), 0, 0, _lib.ptr(ws), nbytes, _lib.stream_ptr())
        return dx, dw, db


def linear_rows(x, weight, bias):
    return LinearRows.apply(x, weight, bias)


def gcn_conv(x, weight, bias, graph, relu=False, p_drop=0.0, site=0, w16=None):
    """``w16``: cached ``split_weight16(weight)`` of the CURRENT weight values (optional)."""
    return GCNConvReLU.apply(x, weight, bias, graph, relu, p_drop, site, w16)


def lstm_head(feats, flat, dims, bw=1, dropout=(0.0, 0.0)):
    return LSTMHead.apply(feats, flat, dims, bw, dropout)
