"""``CellwiseMaskData`` (reference dance/transforms/mask.py:80-291): cell-wise train / valid / test masks for imputation.

The reference loops over cells and calls ``rng.choice(num_positive, n_masked, p=prob, replace=False)`` on a numpy Generator; here
one CUDA block per cell draws the same weighted sample WITHOUT replacement by the Efraimidis-Spirakis keys ``log(u) / w`` with a
counter-based uniform ``u(seed, cell, gene)`` (``b2_cellwise_mask_u8``).  Same distribution and the same per-cell counts
(``floor(n_pos * mask_rate)`` masked, ``max(1, round(0.1 * n))`` of them validation when ``add_test_mask``) — not the same random
stream, so individual masked positions differ from a numpy run with the same seed.  The masks stay on the device until read.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from .base import BaseTransform


class CellwiseMaskData(BaseTransform):
    _DISPLAY_ATTRS = ("distr", "mask_rate", "seed", "min_gene_counts", "add_test_mask")

    def __init__(self, distr: Optional[str] = "exp", mask_rate: Optional[float] = 0.1, seed: Optional[int] = None,
                 min_gene_counts: int = 5, add_test_mask: bool = False, **kwargs):
        super().__init__(**kwargs)
        if not 0.0 <= mask_rate <= 1.0:
            raise ValueError(f"mask_rate must be between 0 and 1, got {mask_rate}")
        if distr not in ("exp", "uniform"):
            raise ValueError(f"Unknown distribution function option {distr!r}, available options are: 'exp', 'uniform'")
        self.distr, self.mask_rate, self.seed = distr, mask_rate, seed
        self.min_gene_counts, self.add_test_mask = min_gene_counts, add_test_mask

    def __call__(self, data):
        adata = data.data
        if not hasattr(adata, "layers"):
            raise AttributeError("Input data object does not have the expected structure 'data.layers'")
        from .filter import _device_matrix
        Xd = _device_matrix(data)
        seed = self.seed if self.seed is not None else int(torch.seed() & 0x7FFFFFFF)     # unseeded: fresh entropy, like default_rng(None)
        train, valid, test, overflow = ops.cellwise_mask(Xd, self.mask_rate, self.min_gene_counts, self.distr, self.add_test_mask, seed)
        if overflow:
            self.logger.warning(f"{overflow} cells have more stored non-zeros than the kernel stages (3072) and were left unmasked")
        adata.layers["train_mask"], adata.layers["valid_mask"], adata.layers["test_mask"] = train, valid, test
        n_total = Xd.numel()
        n_train, n_valid, n_test = int(train.sum()), int(valid.sum()), int(test.sum())
        self.logger.info(f"Masking complete. Total elements: {n_total}")
        self.logger.info(f"  Train mask: {n_train} elements ({n_train/n_total:.2%})")
        self.logger.info(f"  Valid mask: {n_valid} elements ({n_valid/n_total:.4%})")
        if self.add_test_mask:
            self.logger.info(f"  Test mask:  {n_test} elements ({n_test/n_total:.4%})")
        return data
