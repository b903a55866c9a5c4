"""Minimal stand-ins for ``anndata.AnnData`` and ``dance.data.Data`` — just the surface the hot-path
transforms and models touch (reference dance/data/base.py:84-108, 131-168, 203-272, 415-475,
817-851).  When the real ``anndata`` is importable, pass a real AnnData: ``Data`` only uses attribute access.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import scipy.sparse as sp
import torch


class AnnDataLite:
    """Attribute bag with the AnnData field names (X, obs, var, obsm, varm, obsp, varp, layers, uns)."""

    def __init__(self, X, obs=None, var=None, obsm=None, varm=None, obsp=None, varp=None, layers=None, uns=None):
        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else {}
        self.var = var if var is not None else {}
        self.obsm, self.varm = dict(obsm or {}), dict(varm or {})
        self.obsp, self.varp = dict(obsp or {}), dict(varp or {})
        self.layers, self.uns = dict(layers or {}), dict(uns or {})
        self.n_obs, self.n_vars = n, g

    @property
    def shape(self):
        return self.X.shape

    def copy(self):
        import copy
        return copy.deepcopy(self)

    @property
    def var_names(self):
        """Gene identifiers: ``var["names"]`` (or the index of a DataFrame ``var``); defaults to "0".."g-1"."""
        import pandas as pd
        if hasattr(self.var, "index"):
            return pd.Index(self.var.index.astype(str))
        if isinstance(self.var, dict) and "names" in self.var:
            return pd.Index(np.asarray(self.var["names"]).astype(str))
        return pd.Index([str(i) for i in range(self.n_vars)])

    def _inplace_subset_var(self, index):
        """Keep the genes selected by a boolean mask, integer positions or names (AnnData._inplace_subset_var)."""
        index = np.asarray(index)
        if index.dtype == bool:
            keep = np.flatnonzero(index)
        elif index.dtype.kind in "iu":
            keep = index
        else:
            keep = self.var_names.get_indexer(index.astype(str))
            if (keep < 0).any():
                raise KeyError("unknown gene names in _inplace_subset_var")
        self.X = self.X[:, keep]
        if hasattr(self.var, "iloc"):
            self.var = self.var.iloc[keep]
        elif isinstance(self.var, dict):
            self.var = {k: (np.asarray(v)[keep] if np.ndim(v) >= 1 and len(v) == self.n_vars else v) for k, v in self.var.items()}
        self.varm = {k: v[keep] for k, v in self.varm.items()}
        self.varp = {k: v[keep][:, keep] for k, v in self.varp.items()}
        self.layers = {k: v[:, keep] for k, v in self.layers.items()}
        self.n_vars = len(keep)


_CONFIG_KEYS = ("feature_mod", "feature_channel", "feature_channel_type", "label_mod", "label_channel", "label_channel_type")


class Data:
    """``dance.data.Data`` stand-in: holds a reference to the AnnData(-like) object, contiguous
    train | val | test splits, the ``dance_config`` dict and the typed accessors."""

    def __init__(self, data, train_size: Optional[Union[int, str]] = None, val_size: int = 0, test_size: int = -1):
        self._data = data
        if "dance_config" not in data.uns:
            data.uns["dance_config"] = {}
        self._split_idx_dict: Dict[str, List[int]] = {}
        n = data.shape[0]
        if train_size == "all":
            self._split_idx_dict["train"] = list(range(n))
        elif train_size is not None:
            sizes = {"train": train_size, "val": val_size, "test": test_size}
            if list(sizes.values()).count(-1) > 1:
                raise ValueError("Only one split size may be -1")
            known = sum(v for v in sizes.values() if v != -1)
            sizes = {k: (n - known if v == -1 else v) for k, v in sizes.items()}
            start = 0
            for k in ("train", "val", "test"):
                self._split_idx_dict[k] = list(range(start, start + sizes[k]))
                start += sizes[k]

    # -- AnnData mirrors ---------------------------------------------------------
    @property
    def data(self):
        return self._data

    @property
    def config(self) -> Dict[str, Any]:
        return self._data.uns["dance_config"]

    @property
    def shape(self):
        return self._data.shape

    @property
    def num_cells(self):
        return self._data.shape[0]

    @property
    def num_features(self):
        return self._data.shape[1]

    def get_split_idx(self, name: str, error_on_miss: bool = False):
        if name in self._split_idx_dict:
            return self._split_idx_dict[name]
        if error_on_miss:
            raise KeyError(f"Unknown split {name!r}. Please set the split inddices via set_split_idx first.")
        return None

    def set_split_idx(self, name: str, idx: Sequence[int]):
        self._split_idx_dict[name] = list(idx)

    train_idx = property(lambda self: self.get_split_idx("train"))
    val_idx = property(lambda self: self.get_split_idx("val"))
    test_idx = property(lambda self: self.get_split_idx("test"))

    # -- config --------------------------------------------------------------------
    def set_config(self, *, overwrite: bool = False, **kwargs):
        self.set_config_from_dict(kwargs, overwrite=overwrite)

    def set_config_from_dict(self, config_dict: Dict[str, Any], *, overwrite: bool = False):
        for k, v in config_dict.items():
            if k not in _CONFIG_KEYS:
                raise KeyError(f"Unknown config option {k!r}, available options are {_CONFIG_KEYS}")
            if k in self.config and self.config[k] != v and not overwrite:
                raise KeyError(f"Config option {k!r} already set to {self.config[k]!r}; pass overwrite=True to replace it")
            self.config[k] = v

    # -- accessors -------------------------------------------------------------------
    def get_feature(self, *, split_name: Optional[str] = None, return_type: str = "numpy", channel: Optional[str] = None,
                    channel_type: Optional[str] = "obsm", mod: Optional[str] = None):
        if mod is not None:
            raise NotImplementedError("multi-modal (MuData) access is out of scope (SURVEY §2)")
        if channel is None:
            feature = self._data.X
            channel_type = "X"
        else:
            channel_type = channel_type or "obsm"
            if channel_type == "X":
                feature = self._data.X
            else:
                feature = getattr(self._data, channel_type)[channel]
        if return_type == "default":
            if split_name is not None:
                raise ValueError("split_name is not supported when return_type='default'")
            return feature
        if return_type == "sparse":
            feature = sp.csr_matrix(feature)
        else:
            if sp.issparse(feature):
                feature = feature.toarray()
            elif hasattr(feature, "values") and not isinstance(feature, (np.ndarray, torch.Tensor)):
                feature = feature.values
            feature = np.asarray(feature) if not isinstance(feature, torch.Tensor) else feature
        if split_name is not None:
            idx = self.get_split_idx(split_name, error_on_miss=True)
            feature = feature[idx][:, idx] if channel_type == "obsp" else feature[idx]
        if return_type == "torch" and not isinstance(feature, torch.Tensor):
            feature = torch.from_numpy(np.ascontiguousarray(feature))
        return feature

    def _get(self, kind: str, split_name, return_type):
        mods = self.config.get(f"{kind}_mod")
        channels = self.config.get(f"{kind}_channel")
        types = self.config.get(f"{kind}_channel_type")
        if isinstance(channels, list):
            n = len(channels)
            types = types if isinstance(types, list) else [types] * n
            return [self.get_feature(split_name=split_name, return_type=return_type, channel=c, channel_type=t) for c, t in zip(channels, types)]
        return self.get_feature(split_name=split_name, return_type=return_type, channel=channels, channel_type=types, mod=mods)

    def get_x(self, split_name: Optional[str] = None, return_type: str = "numpy"):
        return self._get("feature", split_name, return_type)

    def get_y(self, split_name: Optional[str] = None, return_type: str = "numpy"):
        return self._get("label", split_name, return_type)
