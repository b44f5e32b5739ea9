"""Device-resident region graph: kNN edges -> normalised CSR of A_hat (and A_hat^T).

The reference rebuilds PyG's ``gcn_norm`` inside every ``GCNConv.forward`` (4x per model
forward, model.py:31-40 / hybrid_model.py:65-74).  The graph of a region never changes, so
here it is normalised once (``wf_gcn_norm_csr``) over R = window * N rows and kept on the
device; rows >= N carry only their unit self loop and cost one identity copy.
"""
from __future__ import annotations

import torch

from . import _lib


class RegionGraph:
    """CSR of A_hat by target (forward) and by source (backward) for ONE region."""

    def __init__(self, edge_index, num_rows, device=None):
        if device is None:
            device = edge_index.device if edge_index.is_cuda else torch.device("cuda")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("RegionGraph lives on a CUDA device; there is no CPU fallback")
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
        ei = edge_index.to(device=device, dtype=torch.long).contiguous()
        E = int(ei.shape[1])
        if E > 0:
            lo, hi = int(ei.min()), int(ei.max())
            if lo < 0 or hi >= num_rows:
                raise IndexError(f"edge_index references row {hi if hi >= num_rows else lo}, x has {num_rows} rows")
        self.device, self.R, self.E, self.cap = device, int(num_rows), E, E + int(num_rows)
        self.edge_index = ei
        i32 = dict(dtype=torch.int32, device=device)
        self.rowptr = torch.empty(self.R + 1, **i32)
        self.col = torch.zeros(self.cap, **i32)
        self.val = torch.zeros(self.cap, dtype=torch.float32, device=device)
        self.rowptr_t = torch.empty(self.R + 1, **i32)
        self.col_t = torch.zeros(self.cap, **i32)
        self.val_t = torch.zeros(self.cap, dtype=torch.float32, device=device)
        nbytes = _lib.query("wf_gcn_norm_workspace_bytes", E, self.R)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            _lib.call("wf_gcn_norm_csr", _lib.ptr(ei), E, self.R, _lib.ptr(self.rowptr), _lib.ptr(self.col),
                      _lib.ptr(self.val), _lib.ptr(self.rowptr_t), _lib.ptr(self.col_t), _lib.ptr(self.val_t),
                      _lib.ptr(ws), nbytes, _lib.stream_ptr())
        self._ws = ws  # keep alive until the stream has consumed it
        self._gather_rows = None
        self._agg_rows = None
        _ = self.gather_rows  # computed eagerly (it synchronises): never inside a CUDA-graph capture
        # wf_csr_count_kernel flags an edge outside [0, R) in workspace int[4R] (csrc/wf_graph.cu); the host check above
        # makes that unreachable, but a corrupted graph must not be normalised silently
        if int(ws.view(torch.int32)[4 * self.R].item()) != 0:
            raise IndexError("wf_gcn_norm_csr: edge_index references a row outside the window")
        _ = self.agg_rows

    @property
    def gather_rows(self):
        """i32 list of the rows whose aggregation is NOT the unit self loop (the t = 0 slice for the reference's
        graphs): the only rows the GCN pre-aggregation pass has to touch."""
        if self._gather_rows is None:
            first = self.rowptr[:-1].long()
            deg = (self.rowptr[1:] - self.rowptr[:-1])
            rows = torch.arange(self.R, device=self.device)
            ident = (deg == 1) & (self.col[first].long() == rows) & (self.val[first] == 1.0)
            self._gather_rows = torch.nonzero(~ident).flatten().to(torch.int32).contiguous()
        return self._gather_rows

    @property
    def agg_rows(self):
        """Leading rows of a window that hold every row with neighbours, rounded up to the 128-row tiles of the GEMM
        (0: the graph has no edges besides the self loops)."""
        if self._agg_rows is None:
            g = self.gather_rows
            self._agg_rows = 0 if g.numel() == 0 else min((int(g.max()) + 128) // 128 * 128, (self.R + 127) // 128 * 128)
        return self._agg_rows

    @property
    def nnz(self):
        return int(self.rowptr[-1])


class StackedGraphs:
    """G regions with equal R and capacity, stacked for one task-batched launch."""

    def __init__(self, graphs):
        g0 = graphs[0]
        if any(g.R != g0.R or g.cap != g0.cap for g in graphs):
            raise ValueError("stacked regions must share the row count and edge capacity")
        self.G, self.R, self.cap, self.device = len(graphs), g0.R, g0.cap, g0.device
        self.rowptr = torch.stack([g.rowptr for g in graphs]).contiguous()
        self.col = torch.stack([g.col for g in graphs]).contiguous()
        self.val = torch.stack([g.val for g in graphs]).contiguous()
        self.rowptr_t = torch.stack([g.rowptr_t for g in graphs]).contiguous()
        self.col_t = torch.stack([g.col_t for g in graphs]).contiguous()
        self.val_t = torch.stack([g.val_t for g in graphs]).contiguous()
        lists = [g.gather_rows for g in graphs]
        self.gather_max = max(int(l.numel()) for l in lists)
        self.agg_rows = max(g.agg_rows for g in graphs)
        self.gather_rows = torch.full((self.G, max(self.gather_max, 1)), -1, dtype=torch.int32, device=self.device)
        for i, l in enumerate(lists):
            self.gather_rows[i, :l.numel()] = l

    def assign(self, slot, graph):
        """Overwrite slot ``slot`` with another region's normalised graph, in place (the stacked tensors keep their
        addresses, so captured CUDA graphs that read them stay valid).  Stream-ordered device copies."""
        if graph.R != self.R or graph.cap != self.cap:
            raise ValueError("stacked regions must share the row count and edge capacity")
        if graph.agg_rows > self.agg_rows or int(graph.gather_rows.numel()) > self.gather_rows.shape[1]:
            raise ValueError("the region has more rows with neighbours than the stack was built for")
        for name in ("rowptr", "col", "val", "rowptr_t", "col_t", "val_t"):
            getattr(self, name)[slot].copy_(getattr(graph, name), non_blocking=True)
        self.gather_rows[slot].fill_(-1)
        self.gather_rows[slot, :graph.gather_rows.numel()] = graph.gather_rows

    @property
    def rowptr_stride(self):
        return self.R + 1 if self.G > 1 else 0

    @property
    def csr_stride(self):
        return self.cap if self.G > 1 else 0


_CACHE = {}
_CACHE_MAX = 64


def graph_for(edge_index, num_rows, device):
    """Memoised RegionGraph for an ``edge_index`` tensor as the reference passes it per call."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_rows), str(device))
    hit = _CACHE.get(key)
    if hit is not None and hit[0]() is edge_index:
        return hit[1]
    import weakref

    g = RegionGraph(edge_index, num_rows, device)
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    try:
        _CACHE[key] = (weakref.ref(edge_index), g)
    except TypeError:
        pass
    return g
