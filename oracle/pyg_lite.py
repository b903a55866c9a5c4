"""Restatement of the few torch_geometric 2.4.0 primitives the reference's STAGATE uses (stagate.py:19-20,103,119-125).
TEST INFRASTRUCTURE — torch_geometric is an un-vendored third-party dependency of the reference (install.sh:32), absent
from this image; "parity unpinned" at this boundary: the reference's own GATConv / Stagate code runs on top
of these restated primitives, which follow PyG's documented semantics:

* ``MessagePassing(aggr="add", node_dim=0)``, flow source_to_target: ``j = edge_index[0]`` (source), ``i = edge_index[1]``
  (target).  ``propagate(edge_index, size=None, **kw)`` lifts every ``message`` argument named ``<k>_j`` / ``<k>_i`` from
  ``kw[k]`` (a tensor, or a (source, target) pair), passes ``index = i``, ``ptr = None``, ``size_i = N`` and scatter-adds
  the messages over ``index``.
* ``softmax(src, index, ptr, num_nodes)``: per-target ``exp(src - max_target) / (Σ_target exp + 1e-16)``.
* ``remove_self_loops`` / ``add_self_loops`` on a [2, E] edge index.
"""
from __future__ import annotations

import inspect

import torch


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    shape = (n, ) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    src_max = torch.full(shape, float("-inf"), dtype=src.dtype).scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - src_max.gather(0, idx)).exp()
    out_sum = torch.zeros(shape, dtype=src.dtype).scatter_add(0, idx, out) + 1e-16
    return out / out_sum.gather(0, idx)


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(n, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1), edge_attr


class MessagePassing(torch.nn.Module):

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target" and node_dim == 0, "only what stagate.py uses is restated"

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        n_i = None
        msg_kwargs = {}
        for name in inspect.signature(self.message).parameters:
            if name in ("size_i", "size_j"):
                continue
            if name.endswith("_j") or name.endswith("_i"):
                data = kwargs[name[:-2]]
                side = 0 if name.endswith("_j") else 1
                if isinstance(data, (tuple, list)):
                    data = data[side]
                if data is None:
                    msg_kwargs[name] = None
                    continue
                if side == 1:
                    n_i = data.shape[0]
                msg_kwargs[name] = data.index_select(0, j if side == 0 else i)
        if n_i is None:
            n_i = size[1] if size is not None else int(i.max()) + 1
        if size is not None and size[1] is not None:
            n_i = size[1]
        if "index" in inspect.signature(self.message).parameters:
            msg_kwargs.update(index=i, ptr=None, size_i=n_i)
        out = self.message(**msg_kwargs)
        res = torch.zeros((n_i, ) + tuple(out.shape[1:]), dtype=out.dtype)
        return res.index_add(0, i, out)
