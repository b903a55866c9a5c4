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
from .cell_feature import WeightedFeaturePCA


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


class PCACellFeatureGraph(BaseTransform):
    """WeightedFeaturePCA followed by CellFeatureGraph (cell_feature_graph.py:83-112)."""

    _DISPLAY_ATTRS = ("n_components", "split_name")

    def __init__(self, n_components: int = 400, split_name: Optional[str] = None, *, normalize_edges: bool = True,
                 feat_norm_mode: Optional[str] = None, feat_norm_axis: int = 0, mod: Optional[str] = None, log_level="WARNING"):
        super().__init__(log_level=log_level)
        self.n_components, self.split_name, self.normalize_edges = n_components, split_name, normalize_edges
        self.feat_norm_mode, self.feat_norm_axis, self.mod = feat_norm_mode, feat_norm_axis, mod

    def __call__(self, data):
        WeightedFeaturePCA(self.n_components, self.split_name, feat_norm_mode=self.feat_norm_mode, feat_norm_axis=self.feat_norm_axis,
                           log_level=self.log_level)(data)
        CellFeatureGraph(cell_feature_channel="WeightedFeaturePCA", mod=self.mod, normalize_edges=self.normalize_edges,
                         log_level=self.log_level)(data)
        return data
