"""Oracle (test infrastructure): import the UNMODIFIED reference generator in the build container.

``/root/reference/real_motion_model.py`` cannot be imported as shipped: it needs
``torch_geometric`` (absent, unpinned) and ``pats.data_loading.Skeleton2D`` (package has a
SyntaxError and reads a missing CSV).  This module installs stand-ins in ``sys.modules`` *before*
importing the untouched reference file, and applies decision D1 as a monkey-patched
``UNet1D.forward`` (the shipped forward raises).  Nothing under /root/reference is
edited or copied.

The torch_geometric stand-ins are written edge-list / scatter style (as PyG itself works) on
purpose: oracle/model_oracle.py restates the same layers with dense per-graph adjacency, so the
goldens pin the two formulations against each other.  PyG itself is absent -> that boundary is
"parity unpinned" (DESIGN.md).

Only oracle/make_golden.py and the container-only tests use this; /root/reference does not
exist on the GPU box.
"""
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


class _Linear(nn.Module):
    """Stands in for torch_geometric.nn.dense.linear.Linear (weight [out,in], optional bias)."""

    def __init__(self, cin, cout, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class GATConv(nn.Module):
    """PyG GATConv semantics for (in, out, heads, concat=False) with defaults negative_slope=0.2,
    add_self_loops=True, bias=True, dropout=0; edge_index row0 = source j, row1 = target i."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2):
        super().__init__()
        assert not concat
        self.heads, self.out_channels, self.negative_slope = heads, out_channels, negative_slope
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Linear(in_channels, heads * out_channels, bias=False)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def forward(self, x, edge_index):
        n, H, C = x.shape[0], self.heads, self.out_channels
        h = self.lin(x).view(n, H, C)
        a_src = (h * self.att_src).sum(-1)
        a_dst = (h * self.att_dst).sum(-1)
        keep = edge_index[0] != edge_index[1]                      # remove_self_loops
        loops = torch.arange(n, device=x.device)
        src = torch.cat([edge_index[0][keep], loops])              # add_self_loops
        dst = torch.cat([edge_index[1][keep], loops])
        e = F.leaky_relu(a_src[src] + a_dst[dst], self.negative_slope)        # [E,H]
        e_max = torch.full((n, H), float("-inf"), device=x.device).scatter_reduce(
            0, dst[:, None].expand(-1, H), e, reduce="amax", include_self=True)
        w = torch.exp(e - e_max[dst])
        denom = torch.zeros(n, H, device=x.device).index_add_(0, dst, w)
        alpha = w / denom[dst]
        out = torch.zeros(n, H, C, device=x.device).index_add_(0, dst, alpha[:, :, None] * h[src])
        return out.mean(dim=1) + self.bias


class GraphConv(nn.Module):
    """PyG GraphConv semantics, aggr='add': lin_rel(sum_{j->i} x_j) + lin_root(x_i)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_rel = _Linear(in_channels, out_channels, bias=True)
        self.lin_root = _Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        agg = torch.zeros_like(x).index_add_(0, edge_index[1], x[edge_index[0]])
        return self.lin_rel(agg) + self.lin_root(x)


class _Skeleton2D:
    """Carries only the two constant lists real_motion_model.py reads (:38-40,121-122);
    values restated in oracle/weights.py from pats/data_loading/skeleton.py:94-110,131-148."""

    def __init__(self, *a, **k):
        from .weights import PARENTS
        self.parents = list(PARENTS)
        self.joint_names = ["J%d" % i for i in range(len(PARENTS))]   # names are never used numerically


def _unet_forward_d1(self, x):
    """D1: identical to the shipped UNet1D.forward except up_attention runs before the concat."""
    s0 = self.downsample_layers[0](x)
    x = self.downsample_layers[1](s0)
    s1 = self.downsample_layers[2](x)
    x = self.downsample_layers[3](s1)
    x = self.bottleneck_attention(self.bottleneck(x))
    x = self.up_attention(self.upsample_layers[0](x))
    x = self.upsample_layers[1](torch.cat([x, s1], dim=1))
    x = self.upsample_layers[2](x)
    x = self.upsample_layers[3](torch.cat([x, s0], dim=1))
    return self.final_conv(x)


class _Data:
    """torch_geometric.data.Data as the discriminator uses it (real_motion_model.py:602,614): a bag of x and edge_index."""

    def __init__(self, x=None, edge_index=None):
        self.x, self.edge_index = x, edge_index


class _Batch:
    """torch_geometric.data.Batch.from_data_list: node features concatenated, edge indices offset per graph."""

    def __init__(self, x, edge_index):
        self.x, self.edge_index = x, edge_index

    @classmethod
    def from_data_list(cls, data_list):
        xs, es, off = [], [], 0
        for d in data_list:
            xs.append(d.x)
            es.append(d.edge_index + off)
            off += d.x.shape[0]
        return cls(torch.cat(xs, dim=0), torch.cat(es, dim=1))


def available():
    return os.path.isdir(REFERENCE_ROOT)


def import_reference():
    """Returns the dict of reference modules {mel_features, motion_evaluation, model_layers,
    real_motion_model} imported from /root/reference with the stand-ins installed."""
    if not available():
        raise RuntimeError("/root/reference is not present (it never is on the GPU box)")
    sys.dont_write_bytecode = True           # the tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_nn.GATConv, tg_nn.GraphConv = GATConv, GraphConv
    tg_data.Data, tg_data.Batch = _Data, _Batch               # used by SelfAttention_D.forward only
    tg.nn, tg.data = tg_nn, tg_data
    pats = types.ModuleType("pats")
    pats_dl = types.ModuleType("pats.data_loading")
    pats_dl.Skeleton2D = _Skeleton2D
    pats.data_loading = pats_dl
    saved = {k: sys.modules.get(k) for k in
             ("torch_geometric", "torch_geometric.nn", "torch_geometric.data", "pats", "pats.data_loading")}
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn,
                        "torch_geometric.data": tg_data, "pats": pats, "pats.data_loading": pats_dl})
    try:
        import importlib.util
        mods = {}
        # the product package may have aliased these names (install_dropin); force the reference files
        for name, path in (("mel_features", "pose_video/mel_features.py"),
                           ("motion_evaluation", "motion_evaluation.py"),
                           ("model_layers", "model_layers.py"),
                           ("real_motion_model", "real_motion_model.py")):
            spec = importlib.util.spec_from_file_location("_a2m_ref_" + name, os.path.join(REFERENCE_ROOT, path))
            mod = importlib.util.module_from_spec(spec)
            if name == "model_layers":
                prev = sys.modules.get("model_layers")
                sys.modules["model_layers"] = mod        # real_motion_model does `from model_layers import *`
            spec.loader.exec_module(mod)
            mods[name] = mod
        if prev is None:
            sys.modules.pop("model_layers", None)
        else:
            sys.modules["model_layers"] = prev
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mods["model_layers"].UNet1D.forward = _unet_forward_d1
    return mods
