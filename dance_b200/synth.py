"""Deterministic synthetic single-cell data — ONE generator for the GPU arm, the CPU arm and the dataset
stand-ins.  Every value is a pure function of (seed, global cell index, gene index) through a counter-based hash, so

  * the same cells come out on any device, in any chunking and under any cell sharding (rank r of N generates rows
    [r0, r1) of the SAME matrix the single-GPU run sees);
  * nothing depends on a device RNG stream (``torch._standard_gamma`` has no generator argument — round-1 finding).

Model: counts ~ NegBinomial(mean μ_g·s_c·shift[type_c, g], dispersion 0.5) realised as Poisson(Gamma(2)·mean/2) — the
Gamma(2) draw is the sum of two exponentials, the Poisson draw is CDF inversion for λ < 10 and a rounded Gaussian above —
with gene means μ_g ~ LogNormal(0,1), size factors s_c ~ LogNormal(0,0.5), K latent cell types that multiply 5 % of the
genes by 4, then Bernoulli thinning to the requested density.  Transcendentals differ by an ulp between CPU and CUDA, so
CPU- and GPU-generated matrices agree except for a handful of borderline counts; on one device the output is bit-stable.

This is input synthesis, not part of the hot path: plain torch ops on whichever device is asked for.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Optional, Tuple

import numpy as np
import torch

_M32 = 0xFFFFFFFF


def _hash32(x: torch.Tensor) -> torch.Tensor:
    """lowbias32 integer finaliser on int64 tensors holding values < 2^32 (wrap-around multiplies keep the low 32 bits)."""
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    return x


def _uniform(seed: int, stream: int, rows: torch.Tensor, cols: Optional[torch.Tensor]) -> torch.Tensor:
    """float32 uniforms in (0, 1), one per (row, col) (or per row when cols is None)."""
    h = _hash32(rows + ((seed * 0x9E3779B1 + stream * 0x85EBCA77) & _M32))
    if cols is not None:
        h = _hash32(h[:, None] ^ _hash32(cols + ((stream * 0xC2B2AE3D + 0x27D4EB2F) & _M32))[None, :])
    return ((h >> 8).to(torch.float32) + 0.5) * (1.0 / 16777216.0)


def gene_parameters(n_genes: int, seed: int = 0, n_types: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """(μ_g [G], shift [n_types, G]) — tiny, drawn on the host so that every arm holds identical copies."""
    rng = np.random.default_rng(seed)
    mu = np.exp(rng.normal(size=n_genes)).astype(np.float32)
    shift = np.ones((n_types, n_genes), dtype=np.float32)
    for t in range(n_types):
        shift[t, rng.choice(n_genes, size=max(1, n_genes // 20), replace=False)] = 4.0
    return mu, shift


def cell_types(n: int, seed: int = 0, n_types: int = 10, row_begin: int = 0, device="cpu") -> torch.Tensor:
    rows = torch.arange(row_begin, row_begin + n, dtype=torch.int64, device=device)
    return (_uniform(seed, 2, rows, None) * n_types).long().clamp_(max=n_types - 1)


def _counts(rows: torch.Tensor, n_genes: int, seed: int, mu: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    dev = rows.device
    cols = torch.arange(n_genes, dtype=torch.int64, device=dev)
    u0, u1 = _uniform(seed, 0, rows, None), _uniform(seed, 1, rows, None)
    s_c = torch.exp(0.5 * torch.sqrt(-2.0 * torch.log(u0)) * torch.cos(6.283185307179586 * u1))
    types = (_uniform(seed, 2, rows, None) * shift.shape[0]).long().clamp_(max=shift.shape[0] - 1)
    mean = s_c[:, None] * mu[None, :] * shift[types]
    lam = mean * 0.5 * (-torch.log(_uniform(seed, 3, rows, cols)) - torch.log(_uniform(seed, 4, rows, cols)))
    del mean
    u = _uniform(seed, 5, rows, cols)
    # large λ: rounded Gaussian (Box-Muller with a second uniform)
    z = torch.sqrt(-2.0 * torch.log(u)) * torch.cos(6.283185307179586 * _uniform(seed, 6, rows, cols))
    big = torch.clamp(torch.round(lam + torch.sqrt(lam) * z), min=0.0)
    del z
    # small λ: inversion of the Poisson CDF (P(k > 48 | λ < 10) < 1e-18)
    lam_s = torch.clamp(lam, max=10.0)
    p = torch.exp(-lam_s)
    cdf = p.clone()
    k = torch.zeros_like(lam)
    for i in range(1, 49):
        k += (u > cdf)
        p *= lam_s / i
        cdf += p
    return torch.where(lam < 10.0, k, big)


def pilot_nonzero_fraction(n_genes: int, seed: int = 0, n_types: int = 10, pilot: int = 2048) -> float:
    """Non-zero fraction of the un-thinned counts, estimated ON THE HOST from the first `pilot` cells so that every arm uses
    the identical thinning probability."""
    mu, shift = gene_parameters(n_genes, seed, n_types)
    rows = torch.arange(pilot, dtype=torch.int64)
    c = _counts(rows, n_genes, seed, torch.from_numpy(mu), torch.from_numpy(shift))
    return float((c > 0).float().mean().clamp_min(1e-6))


def expression_counts(n: int, n_genes: int, seed: int = 0, density: float = 0.10, n_types: int = 10, row_begin: int = 0,
                      device="cpu", chunk: int = 32768, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Raw count matrix rows [row_begin, row_begin + n) of the synthetic dataset `seed` as float32 on `device`."""
    dev = torch.device(device)
    mu_h, shift_h = gene_parameters(n_genes, seed, n_types)
    mu, shift = torch.from_numpy(mu_h).to(dev), torch.from_numpy(shift_h).to(dev)
    p_keep = min(1.0, density / pilot_nonzero_fraction(n_genes, seed, n_types))
    X = out if out is not None else torch.empty((n, n_genes), dtype=torch.float32, device=dev)
    cols = torch.arange(n_genes, dtype=torch.int64, device=dev)
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        rows = torch.arange(row_begin + i0, row_begin + i1, dtype=torch.int64, device=dev)
        cnt = _counts(rows, n_genes, seed, mu, shift)
        keep = _uniform(seed, 7, rows, cols) < p_keep
        X[i0:i1] = cnt * keep
    return X


def spatial_coordinates(n: int, seed: int = 0, row_begin: int = 0, device="cpu") -> torch.Tensor:
    """Jittered hex-grid spot coordinates in [0, √N·100]²."""
    rows = torch.arange(row_begin, row_begin + n, dtype=torch.int64, device=device)
    side = int(np.ceil(np.sqrt(max(row_begin + n, 1))))
    gx, gy = (rows % side).float(), (rows // side).float()
    x = (gx + 0.5 * (gy % 2) + 0.3 * (_uniform(seed, 8, rows, None) - 0.5)) * 100.0
    y = (gy * 0.8660254 + 0.3 * (_uniform(seed, 9, rows, None) - 0.5)) * 100.0
    return torch.stack([x, y], 1)


def fingerprint(X: torch.Tensor, head_rows: int = 1024) -> Dict[str, object]:
    """Cheap identity of a (possibly device-resident) matrix: SHA-256 of its first rows, non-zero count and fp64 sum."""
    head = X[:head_rows].detach().cpu().contiguous().numpy()
    return {"sha256_head": hashlib.sha256(head.tobytes()).hexdigest(), "head_rows": int(head.shape[0]),
            "nnz": int((X != 0).sum().item()), "sum": float(X.double().sum().item()), "shape": list(X.shape)}
