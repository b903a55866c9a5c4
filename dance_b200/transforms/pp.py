"""GPU counterparts of the two ``scanpy.pp`` functions the hot-path pipelines call through
``AnnDataTransform`` (reference examples/single_modality/imputation/scgnn2.py:190,
examples/spatial/spatial_domain/spagcn.py pipeline; transforms/normalize.py:563,618-620).
Same call signature for the arguments the reference uses; they mutate ``adata.X`` in place.

The matrix makes one round trip host → HBM → host per call (the AnnData contract keeps X on the host);
``NormalizeTotalLog1P`` fuses both steps into a single kernel pass.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import scipy.sparse as sp
import torch

from .. import ops


def _to_device(adata):
    X = adata.X
    if sp.issparse(X):
        X = X.toarray()
    if not torch.cuda.is_available():
        raise RuntimeError("dance_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).cuda()


def normalize_total(adata, target_sum: Optional[float] = None, exclude_highly_expressed: bool = False, max_fraction: float = 0.05,
                    key_added: Optional[str] = None, layer=None, layers=None, layer_norm=None, inplace: bool = True, copy: bool = False,
                    _log1p: bool = False, _base: Optional[float] = None):
    if layer is not None or layers is not None or layer_norm is not None or copy or not inplace:
        raise NotImplementedError("only in-place normalisation of adata.X is built")
    Xd = _to_device(adata)
    ops.normalize_total_log1p_(Xd, target_sum=target_sum, max_fraction=max_fraction if exclude_highly_expressed else 1.0,
                               normalize=True, log1p=_log1p, base=_base)
    adata.X = Xd.cpu().numpy()


def log1p(adata, base: Optional[float] = None, copy: bool = False, chunked=None, chunk_size=None, layer=None, obsm=None):
    if copy or layer is not None or obsm is not None:
        raise NotImplementedError("only in-place log1p of adata.X is built")
    Xd = _to_device(adata)
    ops.normalize_total_log1p_(Xd, normalize=False, log1p=True, base=base)
    adata.X = Xd.cpu().numpy()
