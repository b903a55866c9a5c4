"""Restatement of the dgl 1.1.3 surface the reference's GraphSCI uses (graphsci.py:14,117-131,255-260): a COO graph
object and ``dgl.nn.GraphConv(norm="both")``.  TEST INFRASTRUCTURE — dgl is an un-vendored third-party dependency
(install.sh:33) absent from this image, so parity is unpinned at this boundary; the reference's own GNNModel / AEModel /
get_loss code runs on top of these primitives (oracle/ref_loader.py::graphsci).

GraphConv.forward (dgl/nn/pytorch/conv/graphconv.py, norm="both", no edge weights):
    feat_src = feat · outdeg.clamp(1)^-0.5 ; if in_feats > out_feats: (feat_src·W) then sum over in-edges, else sum then ·W ;
    rst · indeg.clamp(1)^-0.5 ; + bias ; activation.  Zero-in-degree nodes raise unless allow_zero_in_degree.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _EdgeBatch:
    """What a dgl message UDF receives: ``edges.src[k]`` / ``edges.dst[k]`` (node data gathered per edge), ``edges.data[k]``."""

    def __init__(self, g):
        self.src = {k: v[g.src] for k, v in g.srcdata.items()}
        self.dst = {k: v[g.dst] for k, v in g.dstdata.items()}
        self.data = dict(g.edata)


class _Mean:
    """``dgl.function.mean(msg, out)``: out[v] = mean of the messages on v's in-edges, 0 for in-degree 0."""

    def __init__(self, msg, out):          # called as fn.mean("m", "neigh") and fn.sum(msg="m", out="h")
        self.msg, self.out = msg, out


class _Sum(_Mean):
    """``dgl.function.sum(msg, out)``."""


class function:   # noqa: N801  (mirrors the ``dgl.function`` namespace)
    mean = _Mean
    sum = _Sum


class DGLError(Exception):
    pass


def expand_as_pair(feat, graph=None):
    """dgl.utils.expand_as_pair for homogeneous graphs / full-graph blocks: the same tensor on both sides."""
    return (feat, feat) if not isinstance(feat, tuple) else feat


class Graph:
    """COO multigraph; edge id = position (``dgl.graph((src, dst))``).  Used as a full-graph "block" too: source and destination
    node sets are both all nodes, so ``srcdata`` / ``dstdata`` alias ``ndata`` and ``number_of_dst_nodes() == num_nodes()``."""

    def __init__(self, src, dst, num_nodes=None):
        self.src, self.dst = torch.as_tensor(src).long(), torch.as_tensor(dst).long()
        self.n = int(num_nodes) if num_nodes is not None else int(max(self.src.max(), self.dst.max())) + 1
        self.ndata, self.edata = {}, {}

    # --- the extra surface cell_feature_graph.py:53-69 and models/nn/gnn.py:84-96 touch ---
    @property
    def srcdata(self):
        return self.ndata

    @property
    def dstdata(self):
        return self.ndata

    def number_of_nodes(self):
        return self.n

    def to(self, device):
        return self

    def number_of_dst_nodes(self):
        return self.n

    def nodes(self):
        return torch.arange(self.n)

    def in_edges(self, i, form="uv"):
        eid = torch.nonzero(self.dst == int(i)).flatten()            # ascending edge id, like dgl
        return (self.src[eid], self.dst[eid], eid) if form == "all" else (self.src[eid], self.dst[eid])

    def add_edges(self, u, v, data=None):
        u, v = torch.as_tensor(u).long(), torch.as_tensor(v).long()
        for k, val in self.edata.items():                               # missing keys are zero-filled by dgl
            add = data[k].to(val.dtype) if data and k in data else torch.zeros((u.numel(), ) + tuple(val.shape[1:]), dtype=val.dtype)
            self.edata[k] = torch.cat([val, add])
        self.src, self.dst = torch.cat([self.src, u]), torch.cat([self.dst, v])

    def local_scope(self):
        import contextlib

        @contextlib.contextmanager
        def scope():
            saved_n, saved_e = dict(self.ndata), dict(self.edata)
            try:
                yield
            finally:
                self.ndata, self.edata = saved_n, saved_e

        return scope()

    def update_all(self, message_func, reduce_func):
        msgs = message_func(_EdgeBatch(self))
        m = msgs[reduce_func.msg]
        out = torch.zeros((self.n, ) + tuple(m.shape[1:]), dtype=m.dtype).index_add(0, self.dst, m)
        if not isinstance(reduce_func, _Sum):
            deg = torch.bincount(self.dst, minlength=self.n).clamp(min=1).to(m.dtype)
            out = out / deg.view(-1, *([1] * (m.dim() - 1)))
        self.ndata[reduce_func.out] = out

    def edges(self):
        return self.src.int(), self.dst.int()

    def num_nodes(self):
        return self.n

    def num_edges(self):
        return self.src.numel()

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.n)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.n)


class GraphConv(nn.Module):

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None, allow_zero_in_degree=False):
        super().__init__()
        assert norm in ("both", "right", "none") and weight and bias
        self._in, self._out, self._act, self._allow = in_feats, out_feats, activation, allow_zero_in_degree
        # attribute names of dgl.nn.pytorch.GraphConv that the in-tree WeightedGraphConv subclass reads (graphsc.py:428-484)
        self._norm, self._allow_zero_in_degree, self._activation = norm, allow_zero_in_degree, activation
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        self.bias = nn.Parameter(torch.Tensor(out_feats))
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, graph, feat):
        if not self._allow and (graph.in_degrees() == 0).any():
            raise RuntimeError("There are 0-in-degree nodes in the graph")
        norm_src = graph.out_degrees().to(feat).clamp(min=1).pow(-0.5)
        h = feat * norm_src[:, None]
        agg = lambda x: torch.zeros((graph.n, x.shape[1]), dtype=x.dtype).index_add(0, graph.dst, x[graph.src])
        if self._in > self._out:
            rst = agg(torch.matmul(h, self.weight))
        else:
            rst = torch.matmul(agg(h), self.weight)
        rst = rst * graph.in_degrees().to(feat).clamp(min=1).pow(-0.5)[:, None]
        rst = rst + self.bias
        return self._act(rst) if self._act is not None else rst


# ---------------------------------------------------------------------------------------------------------------
# dgl.dataloading surface used by scdeepsort.py:183,233-236,270-272,321-322: full-neighbourhood blocks over "in" edges
# ---------------------------------------------------------------------------------------------------------------
class Block:
    """Message-flow block of ``NeighborSampler(fanouts=[-1], edge_dir="in")`` for the seed nodes ``dst_nodes``: all their
    in-edges; source nodes = the seeds first (dgl's convention, AdaptiveSAGE relies on it at gnn.py:87), then the remaining
    neighbours; node / edge data are sliced from the parent graph."""

    def __init__(self, g: Graph, dst_nodes):
        dst_nodes = torch.as_tensor(dst_nodes).long()
        pos = torch.full((g.n, ), -1, dtype=torch.long)
        pos[dst_nodes] = torch.arange(dst_nodes.numel())
        eid = torch.nonzero(pos[g.dst] >= 0).flatten()
        extra = torch.unique(g.src[eid])
        extra = extra[pos[extra] < 0]
        self.src_nodes = torch.cat([dst_nodes, extra])
        spos = torch.full((g.n, ), -1, dtype=torch.long)
        spos[self.src_nodes] = torch.arange(self.src_nodes.numel())
        self.dst_nodes = dst_nodes
        self.src, self.dst = spos[g.src[eid]], pos[g.dst[eid]]
        self.srcdata = {k: v[self.src_nodes] for k, v in g.ndata.items()}
        self.dstdata = {k: v[dst_nodes] for k, v in g.ndata.items()}
        self.edata = {k: v[eid] for k, v in g.edata.items()}
        self.n_dst = dst_nodes.numel()

    def to(self, device):
        return self

    def number_of_dst_nodes(self):
        return self.n_dst

    num_dst_nodes = number_of_dst_nodes

    def local_scope(self):
        import contextlib

        @contextlib.contextmanager
        def scope():
            saved = dict(self.srcdata), dict(self.dstdata), dict(self.edata)
            try:
                yield
            finally:
                self.srcdata, self.dstdata, self.edata = saved

        return scope()

    def update_all(self, message_func, reduce_func):
        msgs = message_func(_EdgeBatch(self))
        m = msgs[reduce_func.msg]
        out = torch.zeros((self.n_dst, ) + tuple(m.shape[1:]), dtype=m.dtype).index_add(0, self.dst, m)
        if not isinstance(reduce_func, _Sum):
            deg = torch.bincount(self.dst, minlength=self.n_dst).clamp(min=1).to(m.dtype)
            out = out / deg.view(-1, *([1] * (m.dim() - 1)))
        self.dstdata[reduce_func.out] = out


class NeighborSampler:

    def __init__(self, fanouts, edge_dir="in"):
        assert all(f == -1 for f in fanouts) and edge_dir == "in", "only full in-neighbourhoods are restated"
        self.n_layers = len(fanouts)


class DataLoader:
    """Seeds are visited in ``torch.randperm`` order when ``shuffle`` (global torch RNG) and recorded in ``DataLoader.history``
    so that a fixture can replay exactly the batches the reference saw.  Yields (input_nodes, output_nodes, blocks)."""

    history = []

    def __init__(self, graph, indices, graph_sampler, batch_size=1, shuffle=False, num_workers=0, **kwargs):
        self.g, self.idx, self.sampler, self.bs, self.shuffle = graph, torch.as_tensor(indices).long(), graph_sampler, batch_size, shuffle

    def enable_cpu_affinity(self):
        import contextlib
        return contextlib.nullcontext()

    def __iter__(self):
        order = self.idx[torch.randperm(self.idx.numel())] if self.shuffle else self.idx
        for i in range(0, order.numel(), self.bs):
            seeds = order[i:i + self.bs]
            DataLoader.history.append(seeds.clone())
            blocks, cur = [], seeds
            for _ in range(self.sampler.n_layers):          # outermost layer last in the list, like dgl
                b = Block(self.g, cur)
                blocks.insert(0, b)
                cur = b.src_nodes
            yield cur, seeds, blocks


class TAGConv(torch.nn.Module):
    """``dgl.nn.TAGConv(in_feats, out_feats, k=2, bias=True, activation=None)`` of dgl 1.1.3 (un-vendored): with
    norm = in_degree^-1/2 (clamped to ≥ 1), fstack = [h]; k times: h ← norm · Σ_{u→v} (norm_u · h_u · e_w); then
    Linear(in·(k+1) → out) on the concatenation.  Called by the reference at sctag.py:101-102, 173-174.  Restatement — unpinned."""

    def __init__(self, in_feats, out_feats, k=2, bias=True, activation=None):
        super().__init__()
        self.k, self.activation = k, activation
        self.lin = torch.nn.Linear(in_feats * (k + 1), out_feats, bias=bias)
        torch.nn.init.xavier_normal_(self.lin.weight, gain=torch.nn.init.calculate_gain("relu"))
        if bias:
            torch.nn.init.zeros_(self.lin.bias)

    def forward(self, g, feat, edge_weight=None):
        src, dst = g.edges()
        n = g.num_nodes()
        norm = torch.bincount(dst, minlength=n).float().clamp(min=1).pow(-0.5).unsqueeze(1)
        fstack = [feat]
        for _ in range(self.k):
            h = fstack[-1] * norm
            m = h[src] if edge_weight is None else h[src] * edge_weight.reshape(-1, 1)
            h = torch.zeros(n, feat.shape[1], dtype=feat.dtype).index_add(0, dst, m) * norm
            fstack.append(h)
        out = self.lin(torch.cat(fstack, dim=-1))
        return self.activation(out) if self.activation is not None else out
