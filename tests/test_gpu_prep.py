"""Pre-processing operators upstream of scGNN / GraphSCI on the device: FilterGenesScanpy /
FilterCellsScanpy / FilterGenesTopK against the numpy restatement of the reference formulas (oracle.port), CellwiseMaskData
against the invariants and the sampling distribution of the reference's numpy loop, and the example pipeline of
examples/single_modality/imputation/scgnn2.py:185-196 with X resident on the device between operators."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _counts(n=700, g=300, seed=0, density=0.15):
    from dance_b200 import synth
    return synth.expression_counts(n, g, seed=seed, density=density).numpy()


def test_gene_and_cell_stats_match_numpy(cuda):
    from dance_b200 import ops
    X = _counts(1234, 517)
    Xd = torch.from_numpy(X).to(cuda)
    s, q, k = ops.gene_stats(Xd)
    assert np.allclose(s.cpu().numpy(), X.astype(np.float64).sum(0), rtol=1e-12)
    assert np.allclose(q.cpu().numpy(), (X.astype(np.float64)**2).sum(0), rtol=1e-12)
    assert np.array_equal(k.cpu().numpy(), (X > 0).sum(0))
    cs, ck = ops.cell_stats(Xd)
    assert np.allclose(cs.cpu().numpy(), X.astype(np.float64).sum(1), rtol=1e-12)
    assert np.array_equal(ck.cpu().numpy(), (X > 0).sum(1))
    # padded leading dimension
    wide = torch.zeros(1234, 600, device=cuda)
    wide[:, :517] = Xd
    s2, _, _ = ops.gene_stats(wide[:, :517])
    assert torch.equal(s2, s)
    rows = torch.tensor([5, 3, 1200], device=cuda)
    cols = torch.tensor([516, 0, 7, 7], device=cuda, dtype=torch.int32)
    assert np.array_equal(ops.subset(Xd, rows, cols).cpu().numpy(), X[[5, 3, 1200]][:, [516, 0, 7, 7]])


@pytest.mark.parametrize("kw", [dict(min_cells=0.05), dict(min_cells=30), dict(max_cells=100), dict(min_counts=40), dict(max_counts=500),
                                dict(min_counts=0.3)])
def test_filter_genes_scanpy(cuda, kw):
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import FilterGenesScanpy
    from dance_b200.transforms.filter import get_count
    from oracle import port
    X = _counts()
    data = Data(AnnDataLite(X.copy()))
    FilterGenesScanpy(**kw)(data)
    if "min_counts" in kw and isinstance(kw["min_counts"], float):      # ratio in (0,1): percentile of the gene totals
        ref, _ = port.scanpy_filter(X, "genes", min_counts=np.percentile(X.sum(0), kw["min_counts"] * 100))
    else:
        args = {("min_other" if k == "min_cells" else "max_other" if k == "max_cells" else k): get_count(v, X.shape[0]) if "cells" in k else v
                for k, v in kw.items()}
        ref, _ = port.scanpy_filter(X, "genes", **args)
    assert 0 < ref.sum() < X.shape[1]
    assert np.array_equal(data.data.X, X[:, ref])


@pytest.mark.parametrize("kw", [dict(min_genes=0.1), dict(min_genes=25), dict(max_genes=60), dict(min_counts=150)])
def test_filter_cells_scanpy_updates_splits(cuda, kw):
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import FilterCellsScanpy
    from dance_b200.transforms.filter import get_count
    from oracle import port
    X = _counts()
    data = Data(AnnDataLite(X.copy()), train_size=500)
    FilterCellsScanpy(**kw)(data)
    args = {("min_other" if k == "min_genes" else "max_other" if k == "max_genes" else k): get_count(v, X.shape[1]) if "genes" in k else v
            for k, v in kw.items()}
    ref, _ = port.scanpy_filter(X, "cells", **args)
    assert 0 < ref.sum() < X.shape[0]
    assert np.array_equal(data.data.X, X[ref])
    assert len(data.train_idx) == int(ref[:500].sum()) and len(data.test_idx) == int(ref[500:].sum())
    assert data.train_idx == list(range(len(data.train_idx)))


@pytest.mark.parametrize("mode", ["var", "sum", "cv", "rv"])
def test_filter_genes_topk(cuda, mode):
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import FilterGenesTopK
    from oracle import port
    X = np.log1p(_counts(900, 400, seed=2))
    names = np.array([f"g{(i * 7919) % 400:03d}" for i in range(400)])             # shuffled names: kept genes come out name-sorted
    data = Data(AnnDataLite(X.copy(), var={"names": names}))
    FilterGenesTopK(num_genes=120, mode=mode)(data)
    summary = port.gene_summary(X.astype(np.float64), mode)
    assert np.allclose(data.data.uns["gene_summary"], summary, rtol=1e-9, atol=1e-12)
    ref_mask = port.topk_gene_mask(summary, 120)
    kept = data.data.var_names.tolist()
    assert kept == sorted(names[ref_mask].tolist())
    order = np.argsort(names)
    assert np.array_equal(data.data.X, X[:, order[ref_mask[order]]])


@pytest.mark.parametrize("distr,add_test", [("exp", True), ("uniform", False), ("exp", False)])
def test_cellwise_mask_invariants_and_distribution(cuda, distr, add_test):
    """Per cell: floor(n_pos·rate) masked entries (none when n_pos ≤ min_gene_counts), all of them stored non-zeros, masks disjoint,
    max(1, round(0.1·n)) validation entries with add_test_mask; reproducible for a seed; and the inclusion frequency of a value
    follows the weighted-without-replacement law of numpy's rng.choice(p ∝ exp(−x/20)) — checked against a Monte-Carlo run of that
    very numpy call on one cell."""
    from dance_b200 import ops
    X = _counts(600, 250, seed=5, density=0.3)
    X[3] = 0
    X[4, 10:] = 0                                     # few positives → not masked
    Xd = torch.from_numpy(X).to(cuda)
    tr, va, te, over = ops.cellwise_mask(Xd, 0.1, 5, distr, add_test, seed=11)
    tr, va, te = tr.cpu().numpy(), va.cpu().numpy(), te.cpu().numpy()
    assert over == 0
    n_pos = (X != 0).sum(1)
    exp_masked = np.where(n_pos > 5, np.floor(n_pos * 0.1), 0).astype(int)
    assert np.array_equal((~tr).sum(1), exp_masked)
    assert not ((~tr) & (X == 0)).any()               # only stored non-zeros are masked
    assert np.array_equal(~tr, va | te) and not (va & te).any()
    if add_test:
        nv = np.array([0 if m == 0 else (m if m == 1 else max(1, int(np.round(m * 0.1)))) for m in exp_masked])
        assert np.array_equal(va.sum(1), nv)
    else:
        assert not te.any()
    tr2, _, _, _ = ops.cellwise_mask(Xd, 0.1, 5, distr, add_test, seed=11)
    assert np.array_equal(tr2.cpu().numpy(), tr)
    tr3, _, _, _ = ops.cellwise_mask(Xd, 0.1, 5, distr, add_test, seed=12)
    assert not np.array_equal(tr3.cpu().numpy(), tr)
    # distribution on one cell, replicated: inclusion probability per entry vs numpy's own sampler
    vals = np.array([1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 60, 80, 1, 2, 3, 4, 5, 6, 7, 9], dtype=np.float32)
    reps = 20000
    Xr = torch.from_numpy(np.tile(vals, (reps, 1))).to(cuda)
    trr, _, _, _ = ops.cellwise_mask(Xr, 0.25, 5, distr, False, seed=3)
    freq = 1.0 - trr.float().mean(0).cpu().numpy()
    rng = np.random.default_rng(0)
    p = np.exp(-vals / 20.0) if distr == "exp" else np.ones_like(vals)
    p = p / p.sum()
    hits = np.zeros(len(vals))
    for _ in range(reps):
        hits[rng.choice(len(vals), 5, p=p, replace=False)] += 1
    assert np.abs(freq - hits / reps).max() < 0.015, (freq, hits / reps)


def test_example_pipeline_keeps_x_on_device(cuda):
    """The scGNN example's Compose (scgnn2.py:185-196): filters → top-k genes → masks → log1p → SetConfig, with ONE upload of X."""
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import AnnDataTransform, CellwiseMaskData, Compose, FilterCellsScanpy, FilterGenesScanpy, FilterGenesTopK, SetConfig
    from oracle import port
    X = _counts(1500, 600, seed=7, density=0.1)
    adata = AnnDataLite(X.copy())
    data = Data(adata, train_size=1350)
    pipeline = Compose(FilterGenesScanpy(min_cells=0.01), FilterCellsScanpy(min_genes=0.01), FilterGenesTopK(num_genes=200, mode="var"),
                       CellwiseMaskData(add_test_mask=True, seed=0), AnnDataTransform("scanpy.pp.log1p"),
                       SetConfig({"feature_channel": ["train_mask", "valid_mask", "test_mask"], "feature_channel_type": ["layers"] * 3}))
    pipeline(data)
    assert adata._X_dev is not None and adata._X_host is None          # still resident: nothing has read .X yet
    # oracle: the same steps in numpy
    g_keep, _ = port.scanpy_filter(X, "genes", min_other=int(0.01 * X.shape[0]))
    X1 = X[:, g_keep]
    c_keep, _ = port.scanpy_filter(X1, "cells", min_other=int(0.01 * X1.shape[1]))
    X2 = X1[c_keep]
    mask = port.topk_gene_mask(port.gene_summary(X2.astype(np.float64), "var"), 200)
    names = np.array([str(i) for i in np.flatnonzero(g_keep)])
    order = np.argsort(names)
    X3 = X2[:, order[mask[order]]]
    assert np.allclose(data.data.X, np.log1p(X3), rtol=2e-6, atol=1e-7)
    train_mask, valid_mask, test_mask = data.get_x(return_type="default")
    assert train_mask.shape == X3.shape and train_mask.dtype == bool and (valid_mask | test_mask | train_mask).all()
