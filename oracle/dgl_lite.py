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


class Graph:

    def __init__(self, src, dst, num_nodes):
        self.src, self.dst = torch.as_tensor(src).long(), torch.as_tensor(dst).long()
        self.n = int(num_nodes)
        self.ndata, self.edata = {}, {}

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
        assert norm == "both" and weight and bias
        self._in, self._out, self._act, self._allow = in_feats, out_feats, activation, allow_zero_in_degree
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
