"""Graph-building transforms on the GPU kernels.

``CellFeatureGraph`` keeps the reference's constructor, ``out`` channel (``uns["CellFeatureGraph"]``), node
ordering (genes first), edge order, weights and node data names (reference
dance/transforms/graph/cell_feature_graph.py:12-79, incl. the ``cell_id``/``feat_id`` naming quirk :56-59)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import scipy.sparse as sp
import torch

from .. import ops
from ..graph import GraphLite
from .base import BaseTransform


class CellFeatureGraph(BaseTransform):

    def __init__(self, cell_feature_channel: str, gene_feature_channel: Optional[str] = None, *, mod: Optional[str] = None,
                 normalize_edges: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.cell_feature_channel = cell_feature_channel
        self.gene_feature_channel = gene_feature_channel or cell_feature_channel
        self.mod = mod
        self.normalize_edges = normalize_edges

    def __call__(self, data):
        feat = data.get_feature(return_type="default", mod=self.mod)
        if sp.issparse(feat):
            feat = feat.toarray()
        num_cells, num_feats = feat.shape
        X = torch.as_tensor(np.ascontiguousarray(feat, dtype=np.float32)).cuda()
        src, dst, w, nnz = ops.cellgene_graph(X, self.normalize_edges)
        self.logger.info(f"Number of nonzero entries: {nnz:,}")
        self.logger.info(f"Nonzero rate = {nnz / num_cells / num_feats:.1%}")
        g = GraphLite(src.cpu(), dst.cpu(), num_cells + num_feats)      # stored on the host: the dataset cache pickles it
        g.edata["weight"] = w.cpu()
        g.ndata["cell_id"] = torch.concat((torch.arange(num_feats, dtype=torch.int32), -torch.ones(num_cells, dtype=torch.int32)))
        g.ndata["feat_id"] = torch.concat((-torch.ones(num_feats, dtype=torch.int32), torch.arange(num_cells, dtype=torch.int32)))
        gene_feature = data.get_feature(return_type="torch", channel=self.gene_feature_channel, mod=self.mod, channel_type="varm")
        cell_feature = data.get_feature(return_type="torch", channel=self.cell_feature_channel, mod=self.mod, channel_type="obsm")
        g.ndata["features"] = torch.vstack((gene_feature, cell_feature))
        data.data.uns[self.out] = g
        return data
