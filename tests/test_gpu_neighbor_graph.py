"""NeighborGraph (a5): exact kNN + UMAP fuzzy-simplicial-set connectivities against the loop-by-loop restatement of
scanpy/umap in oracle/port.py (third-party, un-vendored → parity unpinned)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _knn_host(X, k):
    d2 = ((X[:, None, :].astype(np.float64) - X[None, :, :].astype(np.float64))**2).sum(-1)
    idx = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int32), np.sqrt(np.take_along_axis(d2, idx, 1)).astype(np.float32)


@pytest.mark.parametrize("n,d,k", [(400, 10, 15), (257, 3, 5), (90, 50, 30)])
def test_umap_connectivities_match_restatement(cuda, n, d, k):
    from dance_b200 import ops
    from oracle import port
    rng = np.random.default_rng(n)
    X = (rng.normal(size=(n, d)) + rng.integers(0, 3, size=(n, 1)) * 4).astype(np.float32)
    X[7] = X[3]                                   # duplicate cell: zero distance to a non-self neighbour
    idx, dist = _knn_host(X, k)
    ref = port.umap_connectivities(idx, dist)
    out = ops.umap_connectivities(torch.from_numpy(idx).to(cuda), torch.from_numpy(dist).to(cuda))
    assert np.array_equal(out.rowptr.cpu().numpy(), ref.indptr) and np.array_equal(out.colidx.cpu().numpy(), ref.indices)
    assert np.allclose(out.vals.cpu().numpy(), ref.data, rtol=2e-5, atol=1e-7)
    dense = ref.toarray()
    assert np.allclose(dense, dense.T)


def test_neighbor_graph_transform(cuda):
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms import NeighborGraph
    from oracle import port
    rng = np.random.default_rng(5)
    n = 600
    rep = (rng.normal(size=(n, 20)) + rng.integers(0, 4, size=(n, 1)) * 3).astype(np.float32)
    data = Data(AnnDataLite(np.zeros((n, 2), np.float32), obsm={"CellPCA": rep}))
    NeighborGraph(n_neighbors=15)(data)
    adj = data.data.obsp["NeighborGraph"]
    idx, dist = _knn_host(rep, 15)
    ref = port.umap_connectivities(idx, dist)
    assert adj.shape == (n, n) and adj.dtype == np.float32
    assert np.array_equal(adj.indptr, ref.indptr) and np.array_equal(adj.indices, ref.indices)
    assert np.allclose(adj.data, ref.data, rtol=2e-5, atol=1e-7)
    assert repr(NeighborGraph()) == "NeighborGraph(n_neighbors=15, n_pcs=None, knn=True, random_state=0, method='umap', metric='euclidean')"
    with pytest.raises(NotImplementedError):
        NeighborGraph(method="gauss")(data)
