"""``NormalizeTotal``, ``Log1P``, ``NormalizeTotalLog1P`` (reference dance/transforms/normalize.py:531-564,
569-628, 664-679) on the GPU kernels.  Constructor arguments, defaults and ``repr`` follow the reference
(``NormalizeTotal`` always passes ``exclude_highly_expressed=True``, normalize.py:618-620)."""
from __future__ import annotations

from numbers import Number
from typing import Optional

import numpy as np
import scipy.sparse

from . import pp
from .base import BaseTransform
from .interface import AnnDataTransform


class Log1P(AnnDataTransform):

    def __init__(self, base: Optional[Number] = None, copy: bool = False, chunked: bool = None, chunk_size: Optional[int] = None,
                 layer: Optional[str] = None, obsm: Optional[str] = None, **kwargs):
        super().__init__("scanpy.pp.log1p", base=base, chunked=chunked, chunk_size=chunk_size, layer=layer, obsm=obsm, copy=copy,
                         **kwargs)


class NormalizeTotal(AnnDataTransform):

    def __init__(self, target_sum: Optional[float] = None, max_fraction: float = 0.05, key_added: Optional[str] = None,
                 layer: Optional[str] = None, layers=None, layer_norm: Optional[str] = None, inplace: bool = True, copy: bool = False,
                 **kwargs):
        super().__init__("scanpy.pp.normalize_total", target_sum=target_sum, key_added=key_added, layer=layer, layers=layers,
                         layer_norm=layer_norm, inplace=inplace, copy=copy, exclude_highly_expressed=True, max_fraction=max_fraction,
                         **kwargs)
        if max_fraction == 1.0:
            self.logger.info("max_fraction set to 1.0, this is equivalent to setting exclude_highly_expressed=False.")

    def __call__(self, data):
        if scipy.sparse.issparse(data.data.X):
            data.data.X = np.array(data.data.X.todense())
        return super().__call__(data)


class NormalizeTotalLog1P(BaseTransform):
    """Both steps in ONE kernel pass over the matrix (the reference runs them back to back, normalize.py:675-679)."""

    def __init__(self, base=None, target_sum=None, max_fraction=0.05, **kwargs):
        super().__init__(**kwargs)
        self.base, self.target_sum, self.max_fraction = base, target_sum, max_fraction

    def __call__(self, data):
        pp.normalize_total(data.data, target_sum=self.target_sum, exclude_highly_expressed=True, max_fraction=self.max_fraction,
                           _log1p=True, _base=self.base)
        return data
