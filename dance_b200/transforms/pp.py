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
    """Device copy of ``adata.X``: the stand-in AnnData keeps it resident between operators (data.AnnDataLite.device_X); a foreign
    AnnData is uploaded for this call."""
    if hasattr(adata, "device_X"):
        return adata.device_X()
    X = adata.X
    if sp.issparse(X):
        X = X.toarray()
    if not torch.cuda.is_available():
        raise RuntimeError("dance_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).cuda()


def _store(adata, Xd):
    if hasattr(adata, "set_device_X"):
        adata.set_device_X(Xd)          # stays in HBM; the host array is rebuilt on the next `.X` read
    else:
        adata.X = Xd.cpu().numpy()


def normalize_total(adata, target_sum: Optional[float] = None, exclude_highly_expressed: bool = False, max_fraction: float = 0.05,
                    key_added: Optional[str] = None, layer=None, layers=None, layer_norm=None, inplace: bool = True, copy: bool = False,
                    _log1p: bool = False, _base: Optional[float] = None):
    if layer is not None or layers is not None or layer_norm is not None or copy or not inplace:
        raise NotImplementedError("only in-place normalisation of adata.X is built")
    Xd = _to_device(adata)
    ops.normalize_total_log1p_(Xd, target_sum=target_sum, max_fraction=max_fraction if exclude_highly_expressed else 1.0,
                               normalize=True, log1p=_log1p, base=_base)
    _store(adata, Xd)


def log1p(adata, base: Optional[float] = None, copy: bool = False, chunked=None, chunk_size=None, layer=None, obsm=None):
    if copy or layer is not None or obsm is not None:
        raise NotImplementedError("only in-place log1p of adata.X is built")
    Xd = _to_device(adata)
    ops.normalize_total_log1p_(Xd, normalize=False, log1p=True, base=base)
    _store(adata, Xd)


def filter_genes(data, min_counts=None, min_cells=None, max_counts=None, max_cells=None, inplace: bool = True, copy: bool = False):
    """``scanpy.pp.filter_genes``: keep genes by total counts or by the number of cells expressing them (exactly one criterion per
    call, like scanpy).  ``inplace=False`` returns ``(gene_subset, number_per_gene)`` as numpy arrays."""
    return _filter(data, "genes", min_counts, min_cells, max_counts, max_cells, inplace, copy)


def filter_cells(data, min_counts=None, min_genes=None, max_counts=None, max_genes=None, inplace: bool = True, copy: bool = False):
    """``scanpy.pp.filter_cells``."""
    return _filter(data, "cells", min_counts, min_genes, max_counts, max_genes, inplace, copy)


def _filter(data, target, min_counts, min_other, max_counts, max_other, inplace, copy):
    if copy:
        raise NotImplementedError("copy=True is not built")
    given = [o is not None for o in (min_counts, min_other, max_counts, max_other)]
    if sum(given) != 1:
        other = "cells" if target == "genes" else "genes"
        raise ValueError(f"Only provide one of the optional parameters `min_counts`, `min_{other}`, `max_counts`, `max_{other}` per call.")
    is_adata = hasattr(data, "X") and not isinstance(data, (np.ndarray, torch.Tensor))
    if is_adata:
        Xd = _to_device(data)
    elif isinstance(data, torch.Tensor):
        Xd = data
    else:
        Xd = torch.as_tensor(np.ascontiguousarray(data.toarray() if sp.issparse(data) else data, dtype=np.float32)).cuda()
    if target == "genes":
        s, _, k = ops.gene_stats(Xd, want_sumsq=False)
    else:
        s, k = ops.cell_stats(Xd)
    use_counts = min_counts is not None or max_counts is not None
    number = s if use_counts else k
    lo = min_counts if min_counts is not None else min_other
    hi = max_counts if max_counts is not None else max_other
    subset = number >= lo if lo is not None else number <= hi
    subset_np = subset.cpu().numpy()
    number_np = number.cpu().numpy()
    number_np = number_np if use_counts else number_np.astype(np.int64)
    if not inplace or not is_adata:
        return subset_np, number_np
    label = "n_counts" if use_counts else ("n_cells" if target == "genes" else "n_genes")
    if target == "genes":
        data.var[label] = number_np
        data._inplace_subset_var(subset_np)
    else:
        data.obs[label] = number_np
        data._inplace_subset_obs(subset_np)
