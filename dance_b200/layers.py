"""Graph-convolution layers with explicit forward / backward on the C-ABI kernels.

``GraphConvLayer`` covers the two GraphConv flavours on the hot path:

* ``dgl.nn.GraphConv(norm="both")`` as used by GraphSCI (graphsci.py:117-131): no edge weights; multiplies by W first
  when ``in_feats > out_feats`` and aggregates first otherwise;
* the in-tree ``WeightedGraphConv.forward`` of graph-sc (modules/single_modality/clustering/graphsc.py:428-484): messages
  are ``h_src · w_e`` (edge weights), W always first, ``agg`` "sum" or "mean", ``norm`` "both" | "right" | "none".

Both are ``out = act(N_dst · A · (N_src · X [· W]) [· W] + b)`` with diagonal degree scalings, so the layer folds the
scalings into the CSR values once (``bind``) and each call is one SpMM + one GEMM (bias / activation fused).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import scipy.sparse as sp
import torch

from . import ops


class GraphConvLayer:

    def __init__(self, in_feats: int, out_feats: int, *, norm: str = "both", edge_weighted: bool = False, weight_first: Optional[bool] = None,
                 agg: str = "sum", bias: bool = True, activation: Optional[str] = None, device="cuda", precision: Optional[str] = None,
                 seed: Optional[int] = None):
        if norm not in ("both", "right", "none") or agg not in ("sum", "mean"):
            raise ValueError("norm must be both|right|none and agg sum|mean")
        self.in_feats, self.out_feats, self.norm, self.agg = in_feats, out_feats, norm, agg
        self.edge_weighted, self.activation, self.precision = edge_weighted, activation, precision
        self.weight_first = (in_feats > out_feats) if weight_first is None else weight_first
        self.device = torch.device(device)
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        a = (6.0 / (in_feats + out_feats))**0.5                      # xavier_uniform_ (dgl GraphConv.reset_parameters)
        self.weight = ((torch.rand((in_feats, out_feats), generator=gen) * 2 - 1) * a).to(self.device)
        self.bias = torch.zeros(out_feats, device=self.device) if bias else None
        self.grad_weight = torch.zeros_like(self.weight)
        self.grad_bias = torch.zeros_like(self.bias) if bias else None
        self._A = self._AT = None

    # ---- graph ------------------------------------------------------------------------------
    def bind(self, src, dst, num_nodes: int, edge_weight=None, allow_zero_in_degree: bool = False):
        """Edges u→v (messages flow src→dst).  Degrees are structural (edge counts), as in dgl."""
        src, dst = np.asarray(src).astype(np.int64), np.asarray(dst).astype(np.int64)
        n = int(num_nodes)
        indeg = np.bincount(dst, minlength=n)
        outdeg = np.bincount(src, minlength=n)
        if not allow_zero_in_degree and (indeg == 0).any():
            raise RuntimeError("There are 0-in-degree nodes in the graph")
        w = np.ones(len(src), np.float32) if (edge_weight is None or not self.edge_weighted) else np.asarray(edge_weight, np.float32).reshape(-1)
        vals = w.astype(np.float32)
        if self.norm == "both":
            vals = vals * (outdeg.clip(min=1).astype(np.float32)**-0.5)[src]
        if self.norm == "both":
            vals = vals * (indeg.clip(min=1).astype(np.float32)**-0.5)[dst]
        elif self.norm == "right":
            vals = vals / indeg.clip(min=1).astype(np.float32)[dst]
        if self.agg == "mean":
            vals = vals / indeg.clip(min=1).astype(np.float32)[dst]
        A = sp.csr_matrix((vals, (dst, src)), shape=(n, n))           # row = destination; duplicate edges add up, like update_all(sum)
        A.sort_indices()
        self._A = ops.CSR.from_scipy(A, device=self.device)
        self._AT, _ = ops.csr_transpose(self._A)
        return self

    # ---- forward / backward -----------------------------------------------------------------
    def forward(self, feat: torch.Tensor) -> torch.Tensor:
        if self._A is None:
            raise RuntimeError("bind() the graph first")
        self._x = feat
        if self.weight_first:
            self._h = ops.gemm(feat, self.weight, precision=self.precision)
            self._out = ops.spmm(self._A, self._h, act=self.activation, bias=self.bias)
        else:
            self._h = ops.spmm(self._A, feat)
            self._out = ops.gemm(self._h, self.weight, bias=self.bias, act=self.activation, precision=self.precision)
        return self._out

    __call__ = forward

    def backward(self, dout: torch.Tensor, need_input_grad: bool = True) -> Optional[torch.Tensor]:
        """Accumulates nothing: ``grad_weight`` / ``grad_bias`` are overwritten; returns d(feat) or None."""
        act = self.activation
        if act in ("relu", ):
            dpre = ops.relu_bwd(dout, self._out)
        elif act in ("tanh", "elu"):
            dpre, _ = ops.gat_combine_bwd(dout, self._out, 1, self.out_feats, True, act=act)
        elif act is None:
            dpre = dout
        else:
            raise NotImplementedError(act)
        if self.bias is not None:
            ops.colsum(dpre, out=self.grad_bias)
        if self.weight_first:
            dh = ops.spmm(self._AT, dpre)                                                   # d(X·W)
            ops.gemm(self._x, dh, transA=True, out=self.grad_weight, precision=self.precision)
            return ops.gemm(dh, self.weight, transB=True, precision=self.precision) if need_input_grad else None
        ops.gemm(self._h, dpre, transA=True, out=self.grad_weight, precision=self.precision)
        if not need_input_grad:
            return None
        return ops.spmm(self._AT, ops.gemm(dpre, self.weight, transB=True, precision=self.precision))
