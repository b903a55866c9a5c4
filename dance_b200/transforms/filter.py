"""Gene filters on the hot path's input side (reference dance/transforms/filter.py)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

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
