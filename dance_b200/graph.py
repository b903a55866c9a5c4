"""``GraphLite`` — the slice of the DGL graph surface the hot-path callers touch: node/edge
data dicts, degrees, ``add_edges``, node-induced ``subgraph`` (relabelled in the given order, edges kept in
their relative order, DGL semantics restated in SURVEY App. A), ``to(device)``, pickling.  Edge ids are
positions in the ``src``/``dst`` arrays, exactly like ``dgl.graph((src, dst))``."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, Optional

import torch

from . import ops


class GraphLite:

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: Optional[int] = None):
        self.src = src.to(torch.int64)
        self.dst = dst.to(torch.int64)
        self._n = int(num_nodes) if num_nodes is not None else int(max(self.src.max().item(), self.dst.max().item())) + 1
        self.ndata: Dict[str, torch.Tensor] = {}
        self.edata: Dict[str, torch.Tensor] = {}
        self._csr_in = None

    # ---- sizes -------------------------------------------------------------------
    def number_of_nodes(self) -> int:
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self) -> int:
        return self.src.numel()

    num_edges = number_of_edges

    def nodes(self) -> torch.Tensor:
        return torch.arange(self._n, dtype=torch.int64, device=self.src.device)

    def edges(self):
        return self.src, self.dst

    @property
    def device(self):
        return self.src.device

    def in_degrees(self) -> torch.Tensor:
        return torch.bincount(self.dst, minlength=self._n)

    def out_degrees(self) -> torch.Tensor:
        return torch.bincount(self.src, minlength=self._n)

    # ---- mutation ----------------------------------------------------------------------
    def add_edges(self, u, v, data: Optional[Dict[str, torch.Tensor]] = None):
        u, v = torch.as_tensor(u, dtype=torch.int64, device=self.device), torch.as_tensor(v, dtype=torch.int64, device=self.device)
        k = u.numel()
        self.src, self.dst = torch.cat([self.src, u]), torch.cat([self.dst, v])
        for key, val in self.edata.items():   # missing keys are zero-filled (DGL semantics)
            add = data[key].to(val) if data and key in data else torch.zeros((k, ) + tuple(val.shape[1:]), dtype=val.dtype, device=val.device)
            self.edata[key] = torch.cat([val, add])
        self._csr_in = None

    # ---- views -----------------------------------------------------------------------------
    def subgraph(self, nodes) -> "GraphLite":
        """Node-induced subgraph; nodes are relabelled in the order given, edges keep their relative order."""
        nodes = torch.as_tensor(nodes, dtype=torch.int64, device=self.device)
        new_id = torch.full((self._n, ), -1, dtype=torch.int64, device=self.device)
        new_id[nodes] = torch.arange(nodes.numel(), dtype=torch.int64, device=self.device)
        s, d = new_id[self.src], new_id[self.dst]
        keep = (s >= 0) & (d >= 0)
        g = GraphLite(s[keep], d[keep], nodes.numel())
        g.ndata = {k: v[nodes] for k, v in self.ndata.items()}
        g.edata = {k: v[keep] for k, v in self.edata.items()}
        g.ndata["_ID"] = nodes
        g.edata["_ID"] = torch.nonzero(keep, as_tuple=False).flatten()
        return g

    def to(self, device) -> "GraphLite":
        g = GraphLite(self.src.to(device), self.dst.to(device), self._n)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        return g

    @contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield self
        finally:
            self.ndata, self.edata = nd, ed

    # ---- kernels' view -----------------------------------------------------------------------
    def csr_by_destination(self, weight_key: Optional[str] = "weight"):
        """Destination-indexed CSR (row v = sources of v's in-edges) + the edge permutation (CUDA graphs only)."""
        if self._csr_in is None:
            n = self._n
            order = torch.argsort(self.dst, stable=True)
            rowptr = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
            rowptr[1:] = torch.cumsum(torch.bincount(self.dst, minlength=n), 0)
            self._csr_in = (rowptr.to(torch.int32), self.src[order].to(torch.int32).contiguous(), order)
        rowptr, col, order = self._csr_in
        vals = self.edata[weight_key].reshape(-1)[order].contiguous() if weight_key and weight_key in self.edata else None
        return ops.CSR(rowptr, col, vals, (self._n, self._n)), order

    def __getstate__(self):
        return {"src": self.src.cpu(), "dst": self.dst.cpu(), "n": self._n, "ndata": {k: v.cpu() for k, v in self.ndata.items()},
                "edata": {k: v.cpu() for k, v in self.edata.items()}}

    def __setstate__(self, st):
        self.src, self.dst, self._n, self.ndata, self.edata, self._csr_in = st["src"], st["dst"], st["n"], st["ndata"], st["edata"], None
