"""Gene filters on the hot path's input side (reference dance/transforms/filter.py)."""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

from .. import ops
from . import pp
from .base import BaseTransform


class FilterGenesMatch(BaseTransform):
    """Drop genes whose name starts / ends with one of the given prefixes / suffixes (filter.py:386-435, including the
    reference's inverted ``case_sensitive`` switch: ``True`` upper-cases both sides, i.e. matches case-INsensitively)."""

    _DISPLAY_ATTRS = ("prefixes", "suffixes")

    def __init__(self, prefixes: Optional[List[str]] = None, suffixes: Optional[List[str]] = None, case_sensitive: bool = False,
                 **kwargs):
        super().__init__(**kwargs)
        self.prefixes = prefixes or []
        self.suffixes = suffixes or []
        self.case_sensitive = case_sensitive
        if case_sensitive:
            self.prefixes = [i.upper() for i in self.prefixes]
            self.suffixes = [i.upper() for i in self.suffixes]

    def __call__(self, data):
        names = data.data.var_names
        indicator = np.zeros(data.shape[1], dtype=bool)
        for kind, items in (("prefix", self.prefixes), ("suffix", self.suffixes)):
            for item in items:
                ids = names.str.upper().str if self.case_sensitive else names.str
                hit = np.asarray(ids.startswith(item) if kind == "prefix" else ids.endswith(item), dtype=bool)
                self.logger.info(f"{hit.sum()} number of genes will be removed due to {kind} {item!r}")
                indicator |= hit
        self.logger.info(f"Removing {indicator.sum()} genes in total")
        data.data._inplace_subset_var(names[~indicator])
        return data


def get_count(count_or_ratio: Optional[Union[float, int]], total: int) -> Optional[int]:
    """A count, or a ratio of ``total`` (filter.py:28-50)."""
    if count_or_ratio is None:
        return None
    if isinstance(count_or_ratio, float):
        if count_or_ratio > 1.:
            raise ValueError(f"{count_or_ratio=} is greater than 1. Ratio cannot be greater than 1.")
        return int(count_or_ratio * total)
    if isinstance(count_or_ratio, int):
        if count_or_ratio > total:
            raise ValueError(f"{count_or_ratio=} is greater than {total=}")
        return count_or_ratio
    raise TypeError(f"count_or_ratio must be either float or int, got {type(count_or_ratio)}")


def _device_matrix(data, split_name=None, channel=None, channel_type="X"):
    """The matrix an operator works on as a CUDA tensor: resident X when no split / channel is requested."""
    if split_name is None and channel is None and hasattr(data.data, "device_X"):
        return data.data.device_X()
    x = data.get_feature(return_type="numpy", split_name=split_name, channel=channel, channel_type=channel_type)
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()


class FilterScanpy(BaseTransform):
    """``sc.pp.filter_cells`` / ``sc.pp.filter_genes`` with ratios allowed (filter.py:53-158): the per-gene / per-cell counts are one
    device reduction pass (``b2_gene_stats_f32`` / ``b2_cell_stats_f32``), the subsetting a device gather."""

    _FILTER_TARGET = None

    def __init__(self, min_counts=None, min_genes_or_cells=None, max_counts=None, max_genes_or_cells=None, split_name=None, channel=None,
                 channel_type: Optional[str] = "X", key_n_counts=None, key_n_genes_or_cells=None, inplace=True, **kwargs):
        super().__init__(**kwargs)
        self.min_counts, self.max_counts = min_counts, max_counts
        self.min_genes_or_cells, self.max_genes_or_cells = min_genes_or_cells, max_genes_or_cells
        self.split_name, self.channel, self.channel_type = split_name, channel, channel_type
        self.key_n_counts, self.key_n_genes_or_cells, self.inplace = key_n_counts, key_n_genes_or_cells, inplace
        if self._FILTER_TARGET is None:
            raise NotImplementedError("Use FilterCellsScanpy or FilterGenesScanpy instead")
        if self._FILTER_TARGET == "cells":
            self.min_genes, self.max_genes = min_genes_or_cells, max_genes_or_cells
        else:
            self.min_cells, self.max_cells = min_genes_or_cells, max_genes_or_cells

    def _prep_counts(self, totals: torch.Tensor):
        """Fractional min_counts / max_counts in (0, 1) mean percentiles of the totals (filter.py:147-160)."""
        frac = lambda v: isinstance(v, float) and 0 < v < 1
        if frac(self.min_counts) or frac(self.max_counts):
            t = totals.cpu().numpy()
            if isinstance(self.min_counts, float) and 0 <= self.min_counts <= 1:
                return np.percentile(t, self.min_counts * 100), None
            return None, np.percentile(t, self.max_counts * 100)
        return self.min_counts, self.max_counts

    def __call__(self, data):
        Xd = _device_matrix(data, self.split_name, self.channel, self.channel_type)
        total_cells, total_features = Xd.shape
        genes = self._FILTER_TARGET == "genes"
        if genes:
            sums, _, nnz = ops.gene_stats(Xd, want_sumsq=False)
        else:
            sums, nnz = ops.cell_stats(Xd)
        min_counts, max_counts = self._prep_counts(sums)
        basis = total_cells if genes else total_features
        opts = [("counts", min_counts, None), ("other", get_count(self.min_genes_or_cells, basis), None),
                ("counts", None, max_counts), ("other", None, get_count(self.max_genes_or_cells, basis))]
        live = [(k, lo, hi) for k, lo, hi in opts if lo is not None or hi is not None]
        if len(live) != 1:
            other = "cells" if genes else "genes"
            raise ValueError(f"Only provide one of the optional parameters `min_counts`, `min_{other}`, `max_counts`, `max_{other}` per call.")
        kind, lo, hi = live[0]
        number = sums if kind == "counts" else nnz
        subset_ind = (number >= lo if lo is not None else number <= hi).cpu().numpy()
        table = data.data.var if genes else data.data.obs
        if self.key_n_counts is not None:
            self.logger.warning(f"{self.key_n_counts} will be added to the data")
            table[self.key_n_counts] = sums.cpu().numpy()
            if self.key_n_genes_or_cells is not None:
                table[self.key_n_genes_or_cells] = nnz.cpu().numpy().astype(np.int64)
        if not subset_ind.all():
            self.logger.info(f"Subsetting {self._FILTER_TARGET} ({int((~subset_ind).sum()):,} removed) due to {self}")
            if self.inplace:
                if genes:
                    data.data._inplace_subset_var(subset_ind)
                else:
                    data.filter_by_mask(subset_ind)
            else:
                keep = torch.as_tensor(np.flatnonzero(subset_ind), device=Xd.device)
                if genes:
                    data.data.obsm[self.out] = ops.subset(Xd, None, keep.int()).cpu().numpy()
                else:      # the reference stores x[:, subset_ind].T here (filter.py:143) — cells mask applied to columns; kept
                    data.data.varm[self.out] = ops.subset(Xd, None, keep.int()).cpu().numpy().T


class FilterCellsScanpy(FilterScanpy):
    _DISPLAY_ATTRS = ("min_counts", "min_genes", "max_counts", "max_genes", "split_name")
    _FILTER_TARGET = "cells"

    def __init__(self, min_counts=None, min_genes=None, max_counts=None, max_genes=None, split_name=None, channel=None,
                 channel_type: Optional[str] = "X", key_n_counts=None, key_n_genes=None, inplace=True, **kwargs):
        super().__init__(min_counts=min_counts, min_genes_or_cells=min_genes, max_counts=max_counts, max_genes_or_cells=max_genes,
                         split_name=split_name, channel=channel, channel_type=channel_type, key_n_counts=key_n_counts,
                         key_n_genes_or_cells=key_n_genes, inplace=inplace, **kwargs)


class FilterGenesScanpy(FilterScanpy):
    _DISPLAY_ATTRS = ("min_counts", "min_cells", "max_counts", "max_cells", "split_name")
    _FILTER_TARGET = "genes"

    def __init__(self, min_counts=None, min_cells=None, max_counts=None, max_cells=None, split_name=None, channel=None,
                 channel_type: Optional[str] = "X", key_n_counts=None, key_n_cells=None, inplace=True, **kwargs):
        super().__init__(min_counts=min_counts, min_genes_or_cells=min_cells, max_counts=max_counts, max_genes_or_cells=max_cells,
                         split_name=split_name, channel=channel, channel_type=channel_type, key_n_counts=key_n_counts,
                         key_n_genes_or_cells=key_n_cells, inplace=inplace, **kwargs)


_SUMMARY_MODES = ("cv", "rv", "sum", "var")


def gene_summary(Xd: torch.Tensor, mode: str) -> np.ndarray:
    """Per-gene summary statistic of FilterGenes (filter.py:480-489) from one device pass: sum | var (population variance
    E[x²] − E[x]²) | cv (std / mean) | rv (var / mean), non-finite ratios → 0."""
    n = Xd.shape[0]
    s, q, _ = ops.gene_stats(Xd, want_nnz=False)
    s, q = s.cpu().numpy(), q.cpu().numpy()
    if mode == "sum":
        return s
    mean = s / n
    var = q / n - mean * mean
    if mode == "var":
        return var
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (np.sqrt(np.maximum(var, 0)) if mode == "cv" else var) / mean
    return np.nan_to_num(ratio, posinf=0, neginf=0)


class FilterGenes(BaseTransform):
    """Filter genes on a summary of their expression (filter.py:438-520); subclasses pick the genes to keep."""

    def __init__(self, *, mode: str = "sum", channel=None, channel_type=None, whitelist_indicators=None, add_n_counts=True,
                 add_n_cells=True, inplace=True, **kwargs):
        super().__init__(**kwargs)
        if (channel is not None) and (channel_type != "layers"):
            raise ValueError(f"Only X layers is available for filtering genes, specified {channel_type=!r}")
        if mode not in _SUMMARY_MODES:
            raise ValueError(f"Unknown summarization mode {mode!r}, available options are {sorted(_SUMMARY_MODES)}")
        self.mode, self.channel, self.channel_type = mode, channel, channel_type
        self.whitelist_indicators, self.add_n_counts, self.add_n_cells, self.inplace = whitelist_indicators, add_n_counts, add_n_cells, inplace

    def _get_preserve_mask(self, gene_summary: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def __call__(self, data):
        Xd = _device_matrix(data, None, self.channel, self.channel_type)
        if self.add_n_counts or self.add_n_cells:
            s, _, k = ops.gene_stats(Xd, want_sumsq=False)
            if self.add_n_counts:
                self.logger.warning("n_counts will be added to the var of data")
                data.data.var["n_counts"] = s.cpu().numpy()
            if self.add_n_cells:
                self.logger.warning("n_cells will be added to the var of data")
                data.data.var["n_cells"] = k.cpu().numpy().astype(np.int64)
        summary = gene_summary(Xd, self.mode)
        self.logger.info(f"Filtering genes based on {self.mode} expression percentiles in layer {self.channel!r}")
        mask = self._get_preserve_mask(summary)
        names = data.data.var_names
        selected = sorted(names[mask])                    # the reference re-orders the kept genes by NAME (filter.py:492)
        if self.whitelist_indicators is not None:
            cols = [self.whitelist_indicators] if isinstance(self.whitelist_indicators, str) else list(self.whitelist_indicators)
            flag = np.zeros(len(names), dtype=bool)
            for c in cols:
                flag |= np.asarray(data.data.var[c], dtype=bool)
            before = len(selected)
            selected = sorted(set(selected) | set(names[flag]))
            self.logger.info(f"{len(selected) - before:,} genes originally unselected are being added due to whitelist")
        data.data.uns["gene_summary"] = summary
        self.logger.info(f"{data.shape[1] - len(selected):,} genes removed")
        if self.inplace:
            data.data._inplace_subset_var(selected)
        else:
            cols_idx = torch.as_tensor(names.get_indexer(selected), dtype=torch.int32, device=Xd.device)
            data.data.obsm[self.out] = ops.subset(Xd, None, cols_idx).cpu().numpy()


class FilterGenesPercentile(FilterGenes):
    _DISPLAY_ATTRS = ("min_val", "max_val", "mode")

    def __init__(self, min_val: Optional[float] = 1, max_val: Optional[float] = 99, **kwargs):
        kwargs.setdefault("mode", "sum")
        super().__init__(**kwargs)
        self.min_val, self.max_val = min_val, max_val

    def _get_preserve_mask(self, gene_summary):
        lo, hi = np.percentile(gene_summary, self.min_val), np.percentile(gene_summary, self.max_val)
        return np.logical_and(gene_summary >= lo, gene_summary <= hi)


class FilterGenesTopK(FilterGenes):
    """Keep the ``num_genes`` genes with the largest (``top``) / smallest summary value (filter.py:592-664)."""

    _DISPLAY_ATTRS = ("num_genes", "top", "mode")

    def __init__(self, num_genes: int = 1000, top: bool = True, *, mode: str = "cv", channel=None, channel_type: Optional[str] = "X",
                 whitelist_indicators=None, add_n_counts=False, add_n_cells=False, inplace=True, **kwargs):
        if channel is None:
            channel_type = None if channel_type == "X" else channel_type
        super().__init__(mode=mode, channel=channel, channel_type=channel_type, whitelist_indicators=whitelist_indicators,
                         add_n_counts=add_n_counts, add_n_cells=add_n_cells, inplace=inplace, **kwargs)
        self.num_genes, self.top = num_genes, top

    def _get_preserve_mask(self, gene_summary):
        total = gene_summary.size
        if self.num_genes >= total:
            self.logger.warning(f"{self.num_genes=!r} > total number of genes: {total}")
            self.num_genes = total
        order = gene_summary.argsort()
        chosen = order[-self.num_genes:] if self.top else order[:self.num_genes]
        mask = np.zeros(total, dtype=bool)
        mask[chosen] = True
        return mask
