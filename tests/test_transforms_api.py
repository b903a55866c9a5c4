"""Operator API of the transforms boundary: repr / hexdigest / Compose / SetConfig behave like the reference
(tests/transforms/test_basics.py:5-30 style; dance/transforms/base.py:38-45) — no GPU needed — and the
reference's own normalize tests (tests/transforms/test_normalize.py:8-43) re-run on the GPU kernels."""
import numpy as np
import pytest


def test_repr_and_hexdigest():
    from dance_b200.transforms import AnnDataTransform, Compose, Log1P, NormalizeTotal, SetConfig
    t = AnnDataTransform("scanpy.pp.log1p", base=2)
    assert repr(t) == "AnnDataTransform(func=scanpy.pp.log1p, func_kwargs={'base': 2})"   # interface.py:64-66
    s = SetConfig({"feature_channel": "x"})
    assert repr(s) == "SetConfig(config_dict={'feature_channel': 'x'})"
    c = Compose(t, s)
    assert repr(c) == f"Compose(\n  {t!r},\n  {s!r},\n)" and c[1] is s
    assert len(c.hexdigest()) == 32 and c.hexdigest() != Compose(s, t).hexdigest()
    assert NormalizeTotal(target_sum=30, max_fraction=0.99).func_kwargs["exclude_highly_expressed"] is True
    assert Log1P().name == "Log1P"
    # reference tests/transforms/test_basics.py:5-22
    from dance_b200.transforms import CellPCA, PCACellFeatureGraph, WeightedFeaturePCA
    assert repr(CellPCA(n_components=100)) == "CellPCA(n_components=100)"
    assert repr(WeightedFeaturePCA(n_components=100, split_name="train")) == (
        "WeightedFeaturePCA(n_components=100, split_name='train', feat_norm_mode=None, feat_norm_axis=0)")
    assert repr(PCACellFeatureGraph(n_components=100, split_name="train")) == "PCACellFeatureGraph(n_components=100, split_name='train')"
    with pytest.raises(TypeError):
        Compose(t, "not a transform")
    # the remaining graph / filter transforms of the hot path: repr strings are cache keys (datasets/base.py:129-133)
    from dance_b200.transforms import (FeatureFeatureGraph, FilterGenesMatch, NeighborGraph, SpaGCNGraph, SpaGCNGraph2D,
                                       StagateGraph)
    assert repr(NeighborGraph(n_neighbors=10, n_pcs=None, knn=True, random_state=0, method="umap", metric="euclidean")) == (
        "NeighborGraph(n_neighbors=10, n_pcs=None, knn=True, random_state=0, method='umap', metric='euclidean')")   # test_basics.py:17-20
    assert repr(SpaGCNGraph(alpha=1, beta=2)) == "SpaGCNGraph(alpha=1, beta=2)"                                     # test_basics.py:30-32
    assert repr(SpaGCNGraph2D()) == "SpaGCNGraph2D()"
    assert repr(StagateGraph("radius", radius=150)) == "StagateGraph(model_name='radius', radius=150, n_neighbors=5)"
    assert repr(FeatureFeatureGraph(threshold=0.3)) == (
        "FeatureFeatureGraph(threshold=0.3, positive_only=False, normalize_edges=True, score_func='pearson', score_func_kwargs={})")
    assert repr(FilterGenesMatch(prefixes=["ERCC", "MT-"])) == "FilterGenesMatch(prefixes=['ERCC', 'MT-'], suffixes=[])"
    with pytest.raises(ValueError):
        StagateGraph("delaunay")


def test_model_shells_expose_the_reference_surface():
    """Constructor / method names of the reference models — importable and inspectable without a GPU."""
    import inspect
    from dance_b200.modules import graphsci, scdeepsort, scgnn2, spagcn, stagate
    assert list(inspect.signature(spagcn.SpaGCN.fit).parameters)[:3] == ["self", "x", "y"]
    for name in ("search_l", "set_l", "calc_adj_exp", "fit", "predict_proba", "predict", "fit_predict", "score", "preprocessing_pipeline"):
        assert hasattr(spagcn.SpaGCN, name), name
    for name in ("forward", "pretrain", "save_pretrained", "load_pretrained", "fit", "predict", "fit_score", "preprocessing_pipeline"):
        assert hasattr(stagate.Stagate, name), name
    assert list(inspect.signature(graphsci.GraphSCI.__init__).parameters)[1:7] == ["num_cells", "num_genes", "dataset", "dropout", "gpu", "seed"]
    assert list(inspect.signature(graphsci.GraphSCI.fit).parameters)[1:12] == [
        "train_data", "train_data_raw", "graph", "mask", "le", "la", "ke", "ka", "n_epochs", "lr", "weight_decay"]
    assert list(inspect.signature(scdeepsort.ScDeepSort.__init__).parameters)[1:6] == ["dim_in", "dim_hid", "num_layers", "species", "tissue"]
    assert hasattr(scgnn2, "feature_AE_handler") and hasattr(scgnn2, "graph_AE_handler") and hasattr(scgnn2.ScGNN2, "fit")
    # no CPU path: constructing a model on a machine without CUDA fails loudly instead of falling back
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            stagate.Stagate([10, 8, 4])


def test_data_standin_splits_and_config():
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import SetConfig
    ad = AnnDataLite(np.arange(20, dtype=np.float32).reshape(10, 2), layers={"train_mask": np.ones((10, 2), bool)})
    d = Data(ad, train_size=6, val_size=1)
    assert d.train_idx == list(range(6)) and d.val_idx == [6] and d.test_idx == [7, 8, 9]
    SetConfig({"feature_channel": ["train_mask"], "feature_channel_type": ["layers"]})(d)
    (m, ) = d.get_x(return_type="default")
    assert m.shape == (10, 2)
    assert d.get_feature(split_name="test", return_type="numpy").shape == (3, 2)
    with pytest.raises(KeyError):
        d.set_config(feature_channel="other")


@pytest.mark.gpu
def test_normalize_total_reference_tests_on_gpu(cuda, assert_ary_isclose):
    # reference tests/transforms/test_normalize.py:8-30
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import NormalizeTotal
    data = Data(AnnDataLite(X=np.array([[1, 1, 1], [1, 1, 1], [3, 0, 0]], dtype=np.float32)))
    NormalizeTotal(max_fraction=0.99, target_sum=30)(data)
    assert_ary_isclose(data.data.X, np.array([[15.0, 15.0, 15.0], [15.0, 15.0, 15.0], [3.0, 0.0, 0.0]]))
    NormalizeTotal(max_fraction=1.0, target_sum=30)(data)
    assert_ary_isclose(data.data.X, np.array([[10.0, 10.0, 10.0], [10.0, 10.0, 10.0], [30.0, 0.0, 0.0]]))


@pytest.mark.gpu
def test_log1p_and_fused_pipeline_on_gpu(cuda, assert_ary_isclose):
    # reference tests/transforms/test_normalize.py:33-43 and test_interface.py:20-50 (wrapper ≡ direct call)
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import AnnDataTransform, Compose, Log1P, NormalizeTotalLog1P
    from oracle import port
    x = np.array([[1, 1, 1], [1, 1, 1], [3, 0, 0]], dtype=np.float32)
    data = Data(AnnDataLite(X=x.copy()))
    Log1P()(data)
    assert data.data.X.shape == x.shape
    assert_ary_isclose(data.data.X, np.log1p(x))
    X = port.synthetic_expression(200, 150, density=0.2, seed=3, log_normalize=False)
    a, b = Data(AnnDataLite(X=X.copy())), Data(AnnDataLite(X=X.copy()))
    Compose(AnnDataTransform("scanpy.pp.normalize_total", target_sum=1e4), AnnDataTransform("scanpy.pp.log1p"))(a)
    NormalizeTotalLog1P(target_sum=1e4, max_fraction=1.0)(b)
    ref = port.log1p(port.normalize_total(X, target_sum=1e4))
    assert np.allclose(a.data.X, ref, rtol=2e-6, atol=1e-7) and np.allclose(b.data.X, ref, rtol=2e-6, atol=1e-7)


def test_feature_graph_score_helpers_on_cpu():
    """The device-agnostic pieces of FeatureFeatureGraph's spearman / rbf scores against scipy / the reference formulas
    (transforms/graph/feature_feature_graph.py:50-57, utils/matrix.py:70-97); the Gram / correlation kernels are GPU-tested."""
    import torch
    from scipy.stats import rankdata, spearmanr
    from dance_b200.transforms.graph import _average_ranks, _rbf_from_gram
    rng = np.random.default_rng(0)
    X = rng.poisson(0.8, size=(300, 70)).astype(np.float32)           # count data: heavy ties
    X[:, 3] = rng.normal(size=300).astype(np.float32)                 # a tie-free column
    ranks = _average_ranks(torch.from_numpy(X), chunk=16).numpy()
    assert np.array_equal(ranks, rankdata(X, method="average", axis=0).astype(np.float32))
    rho = np.corrcoef(ranks.astype(np.float64), rowvar=False)
    assert np.allclose(rho, spearmanr(X, axis=0)[0], atol=1e-12, equal_nan=True)
    feat = rng.normal(size=(200, 40)).astype(np.float32)
    gram = feat.T.astype(np.float64) @ feat.astype(np.float64)
    nv = np.power(feat.astype(np.float64), 2).sum(0, keepdims=True)
    dist = np.sqrt((nv + nv.T - 2 * gram).clip(0))
    for mode, kw in (("med_dist", {}), ("ind_med_dist", {"denom_scale": 2.0}), ("scale", {"denom_scale": 3.0})):
        denom = {"med_dist": np.median(dist) * kw.get("denom_scale", 1.0), "ind_med_dist": np.median(dist, axis=1, keepdims=True) * kw.get("denom_scale", 1.0),
                 "scale": kw.get("denom_scale", 1.0)}[mode]
        ref = np.exp(-dist / denom)
        out = _rbf_from_gram(torch.from_numpy(gram), scale_mode=mode, **kw).numpy()
        assert np.allclose(out, ref, rtol=1e-9, atol=1e-9), mode
    with pytest.raises(ValueError):
        _rbf_from_gram(torch.eye(3), scale_mode="nope")


def test_spagcn_refine_matches_reference_logic():
    """``refine`` (spagcn.py:290-334): the torch implementation (run here on CPU tensors) against a restatement of the
    reference's pandas loop, on a jittered hexagonal layout with noisy domain labels."""
    import pandas as pd
    import torch
    from dance_b200.modules.spagcn import _refine_labels, refine
    rng = np.random.default_rng(2)
    side = 18
    gx, gy = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    xy = np.stack([gx.ravel() + 0.5 * (gy.ravel() % 2), gy.ravel() * 0.866], 1) + rng.normal(scale=0.01, size=(side * side, 2))
    n = len(xy)
    pred = (xy[:, 0] > side / 2).astype(int) + 2 * (xy[:, 1] > side * 0.433).astype(int)
    noisy = pred.copy()
    flip = rng.random(n) < 0.15
    noisy[flip] = rng.integers(0, 4, flip.sum())
    dis = np.sqrt(((xy[:, None, :] - xy[None, :, :])**2).sum(-1)).astype(np.float32)
    for shape, num_nbs in (("hexagon", 6), ("square", 4)):
        ids = [f"s{i}" for i in range(n)]
        p_df = pd.DataFrame({"pred": noisy}, index=ids)
        d_df = pd.DataFrame(dis, index=ids, columns=ids)
        ref = []
        for i in range(n):                                    # the reference loop, with a stable sort
            nbs = d_df.loc[ids[i], :].sort_values(kind="stable")[0:num_nbs + 1]
            v_c = p_df.loc[nbs.index, "pred"].value_counts()
            self_pred = p_df.loc[ids[i], "pred"]
            ref.append(v_c.idxmax() if (v_c.loc[self_pred] < num_nbs / 2) and (np.max(v_c) > num_nbs / 2) else self_pred)
        # refine() itself always runs on the device; its device-agnostic core is exercised here on CPU tensors
        got = _refine_labels(torch.from_numpy(noisy.astype(np.int64)), torch.from_numpy(dis), num_nbs).tolist()
        assert got == [int(v) for v in ref]
        assert (np.array(got) == pred).mean() > (noisy == pred).mean()       # the vote removes most of the label noise
    out = _refine_labels(torch.tensor([0, 0, 1]), torch.tensor([[0., 1, 2], [1, 0, 1], [2, 1, 0]]), 6)
    assert out.tolist() == [0, 0, 1]                                         # three spots: nobody reaches more than 3 votes
    with pytest.raises(ValueError):
        refine(["a"], [0], np.zeros((1, 1), np.float32), shape="triangle")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            refine(ids, noisy, dis)                                          # fails loudly without a CUDA device
