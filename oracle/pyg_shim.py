"""Stand-ins for the third-party packages the reference imports but this image lacks.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used only to import the
*unmodified* reference files from /root/reference inside the build container,
so that golden vectors can be generated from the reference's own code
(oracle/make_golden.py).  It never travels into the product path.

Restated third-party symbols (torch_geometric is not vendored by the reference
and not pinned in its requirements.txt; PyG >= 2.0 default semantics):

* ``torch_geometric.nn.GCNConv``      -- call sites model.py:4,23-26,31-40,
                                         hybrid_model.py:65-74
* ``torch_geometric.data.Data``       -- dataset.py:3,50-54
* ``torch_geometric.loader.DataLoader`` -- train_hybrid_maml_v5.py:9,121,162,
                                         adapt_hybrid_v5.py:5,182,217
* ``xarray``                          -- only the name ``xr.Dataset`` is touched at
                                         import time (embed_utils.py:10)
"""
import importlib.machinery
import math
import sys
import types

import torch
import torch.nn as nn
import torch.utils.data


def gcn_norm(edge_index, num_nodes, dtype=torch.float32):
    """PyG ``gcn_norm`` with add_self_loops=True, improved=False, flow source->target.

    Existing self loops are dropped, one (i, i) edge of weight 1 is appended for
    every node, the degree is the in-degree by *target* (edge_index[1]), and the
    edge weight is deg[src]^-1/2 * 1 * deg[dst]^-1/2 (inf -> 0).
    """
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    row = torch.cat([row[keep], loops])
    col = torch.cat([col[keep], loops])
    w = torch.ones(row.numel(), dtype=dtype, device=edge_index.device)
    deg = torch.zeros(num_nodes, dtype=dtype, device=edge_index.device)
    deg.scatter_add_(0, col, w)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[row] * w * dis[col]
    return torch.stack([row, col]), w


class _Lin(nn.Module):
    """Bias-free linear holder so the state_dict key is ``<conv>.lin.weight``."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        a = math.sqrt(6.0 / (cin + cout))  # PyG 'glorot'
        with torch.no_grad():
            self.weight.uniform_(-a, a)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight)


class GCNConv(nn.Module):
    """out = scatter_add_{e: dst_e = i}( w_e * (x W^T)[src_e] ) + bias."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        # PyG registers ``bias`` before ``lin`` -> state_dict order bias, lin.weight
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Lin(in_channels, out_channels)

    def forward(self, x, edge_index):
        ei, w = gcn_norm(edge_index, x.size(0), x.dtype)
        h = self.lin(x)
        msg = w.view(-1, 1) * h.index_select(0, ei[0])
        out = torch.zeros_like(h).index_add_(0, ei[1], msg)
        return out + self.bias


class Data:
    def __init__(self, x=None, edge_index=None, y=None):
        self.x, self.edge_index, self.y = x, edge_index, y

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device):
        return Data(self.x.to(device), self.edge_index.to(device), self.y.to(device))


def _collate(items):
    if len(items) == 1:
        return items[0]
    xs, eis, ys, off = [], [], [], 0
    for d in items:  # PyG Batch: node features concatenated, edge ids offset
        xs.append(d.x)
        ys.append(d.y)
        eis.append(d.edge_index + off)
        off += d.num_nodes
    return Data(torch.cat(xs), torch.cat(eis, dim=1), torch.cat(ys))


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, collate_fn=_collate, **kw)


def _module(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)  # torch._dynamo probes find_spec()
    return m


def install():
    """Register the stand-in modules in ``sys.modules`` (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_wf_shim", False):
        return
    tg = _module("torch_geometric")
    tg._wf_shim = True
    tg_nn = _module("torch_geometric.nn")
    tg_nn.GCNConv = GCNConv
    tg_data = _module("torch_geometric.data")
    tg_data.Data = Data
    tg_loader = _module("torch_geometric.loader")
    tg_loader.DataLoader = DataLoader
    tg.nn, tg.data, tg.loader = tg_nn, tg_data, tg_loader
    sys.modules.update({
        "torch_geometric": tg, "torch_geometric.nn": tg_nn,
        "torch_geometric.data": tg_data, "torch_geometric.loader": tg_loader,
    })
    if "xarray" not in sys.modules:
        xr = _module("xarray")
        xr.Dataset = type("Dataset", (), {})
        xr.open_dataset = None
        sys.modules["xarray"] = xr
