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
