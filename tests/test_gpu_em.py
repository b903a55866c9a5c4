"""scGNN EM-iteration stages on the device against the REFERENCE's own functions (fixture
tests/golden/scgnn_em.npz, written by oracle/make_golden_em.py through oracle.ref_loader) and against scikit-learn's KMeans."""
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _adj(g):
    n = len(g["labels"])
    return sp.csr_matrix((np.ones(len(g["adj_indices"]), dtype=np.float32), g["adj_indices"], g["adj_indptr"]), shape=(n, n))


def _state(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")}


def test_sparse_regulariser_weights_equal_dense_column_sums(cuda, golden):
    """graph_celltype_regu_handler + the per-cluster slicing of cluster_AE_handler (scgnn2.py:716-752, 844-846): the column sums of
    the reference's two dense N×N matrices inside each cluster, from the degrees alone (the fixture holds the sums of the matrices
    the reference's own function returned — including its np.matrix product, adjdense[i, j] = deg_j / deg_i)."""
    from dance_b200 import ops
    from dance_b200.modules import scgnn2 as mod
    g = golden("scgnn_em")
    adj = _adj(g)
    A = ops.CSR.from_scipy((adj + sp.eye(adj.shape[0])).tocsr(), cuda, with_values=False)       # the A + I pattern the handlers carry
    w_graph, w_ct = mod.graph_celltype_regu_handler(A, g["labels"])
    assert np.allclose(w_graph.cpu().numpy(), g["w_graph"], rtol=1e-6, atol=1e-7)
    assert np.allclose(w_ct.cpu().numpy(), g["w_celltype"], rtol=1e-6)
    # scipy adjacency without diagonal (what the numpy-returning graph_AE_handler hands out) gives the same
    w2, _ = mod.graph_celltype_regu_handler(adj, g["labels"], cuda)
    assert torch.equal(w2, w_graph)


def test_celltype_loss_gradient_and_l1_match_reference_autograd(cuda, golden):
    """loss_function_graph(regularizer_type="Celltype") + `loss + 1*l1` (scgnn2.py:1268-1274, 1316-1326): value, d/d recon and every
    parameter gradient of one Cluster-AE batch against the reference's autograd."""
    from dance_b200 import ops
    from dance_b200.engine import FeatureAEEngine
    g = golden("scgnn_em")
    members = g["batch_members"]
    x = torch.from_numpy(g["X_recon"][members]).to(cuda)
    xd = torch.from_numpy(g["x_dropout"][members]).to(cuda)
    roww = torch.from_numpy((0.3 + 0.3 * g["w_graph"][members] + 0.1 * g["w_celltype"][members]).astype(np.float32)).to(cuda)
    # the loss kernel alone on the reference's reconstruction
    recon_ref = torch.from_numpy(g["batch_recon"]).to(cuda)
    loss, grad = ops.celltype_loss_grad(recon_ref, x, xd, roww, relu_mask=False)      # gradient w.r.t. the (post-ReLU) reconstruction
    assert abs(loss.item() - float(g["batch_loss"])) < 1e-5 * abs(float(g["batch_loss"]))
    assert rel_err(grad, g["batch_grad_recon"]) < 1e-5
    _, gm = ops.celltype_loss_grad(recon_ref, x, xd, roww, relu_mask=True)              # … and chained through the final ReLU
    assert torch.equal(gm, grad * (recon_ref > 0))
    # the full step: forward, loss, backward, L1 (gradients are read before they are consumed by Adam — the buffer survives the step)
    eng = FeatureAEEngine(x.shape[1], device=cuda, lr=1e-3, precision="fp32")
    eng.load_state_dict(_state(g))
    eng.loss_acc.zero_()
    _, r = eng.train_step(x, None, 0.9, "Celltype", row_weight=roww, x_dropout=xd)
    assert rel_err(r, g["batch_recon"]) < 1e-5
    assert abs(eng.loss_acc.item() - (float(g["batch_loss"]) + float(g["batch_l1"]))) < 1e-5 * abs(float(g["batch_loss"]) + float(g["batch_l1"]))
    for name, gr in eng.params.g.items():
        assert rel_err(gr, g["batch_grad_" + name]) < 1e-4, name


@pytest.mark.parametrize("batch", [12800, 50])
def test_cluster_ae_handler_matches_reference(cuda, golden, batch):
    """cluster_AE_handler end to end (scgnn2.py:821-880): three Cluster-AEs, 3 Adam epochs each, stitched reconstruction.
    batch=50 < cluster size exercises the mini-batch generalisation (full-cluster regulariser weights, per-batch norm term):
    it must stay close to — not equal — the single-batch result."""
    from dance_b200 import ops
    from dance_b200.modules import scgnn2 as mod
    g = golden("scgnn_em")
    adj = _adj(g)
    labels = g["labels"]
    _, lists = mod.cluster_output_handler(labels.tolist())
    args = SimpleNamespace(cluster_AE_batch_size=batch, cluster_AE_epoch=int(g["cluster_epochs"]), cluster_AE_learning_rate=1e-3,
                           cluster_AE_regu_strength=0.9, cluster_AE_dropout_prob=0)
    param = {"device": cuda, "epoch_num": 1, "total_epoch": 2, "n_feature_orig": adj.shape[0] and g["X_recon"].shape[1],
             "x_dropout": g["x_dropout"], "precision": "fp32"}
    param["impute_regu"] = mod.graph_celltype_regu_handler(adj, labels, cuda)
    out = mod.cluster_AE_handler(g["X_recon"], np.zeros_like(g["X_recon"]), lists, args, param, {"model": _state(g)})
    assert isinstance(out, np.ndarray) and out.shape == g["cluster_recon"].shape
    if batch >= 12800:
        assert rel_err(out, g["cluster_recon"]) < 2e-4
    else:      # more (smaller) Adam steps: a different but finite, non-negative reconstruction of the same cells
        assert np.isfinite(out).all() and (out >= 0).all() and not np.allclose(out, g["cluster_recon"])


@pytest.mark.parametrize("n,d,k", [(3000, 16, 4), (20000, 16, 9), (1500, 144, 3)])
def test_device_kmeans_matches_sklearn(cuda, n, d, k):
    """KMeans(n_clusters=k, n_init="auto", random_state=0).fit_predict(embed) (scgnn2.py:186): sklearn's own k-means++ seeding on
    the host, Lloyd iterations on the device — identical labels on separated clusters, equal inertia otherwise."""
    from sklearn.cluster import KMeans
    from dance_b200.modules import scgnn2 as mod
    rng = np.random.default_rng(n)
    X = (rng.normal(size=(n, d)) + rng.integers(0, k + 1, size=(n, 1)) * 2.5).astype(np.float32)
    ref = KMeans(n_clusters=k, n_init="auto", random_state=0).fit(X)
    lab = mod.kmeans_fit_predict(X, k, cuda).cpu().numpy()
    agree = (lab == ref.labels_).mean()
    Xc = X - X.mean(0)
    inertia = sum(((Xc[lab == c] - Xc[lab == c].mean(0))**2).sum() for c in range(k))
    assert agree > 0.999, agree
    assert abs(inertia - ref.inertia_) < 1e-3 * ref.inertia_


def test_clustering_handler_and_full_em_fit(cuda):
    """ScGNN2.fit at the example's structure with EM iterations (total_epoch = 2): Louvain (host C++) → k, KMeans → labels, sparse
    regulariser, Cluster-AEs, Feature-AE / Graph-AE rounds; returns a finite imputed matrix of the input's shape."""
    from dance_b200.modules import scgnn2 as mod
    from oracle import port
    n, g = 900, 120
    X = port.synthetic_expression(n, g, density=0.25, seed=2)
    args = SimpleNamespace(total_epoch=2, feature_AE_batch_size=12800, feature_AE_epoch=[6, 3], feature_AE_learning_rate=1e-3,
                           feature_AE_regu_strength=0.9, feature_AE_dropout_prob=0, feature_AE_concat_prev_embed=None,
                           graph_AE_epoch=5, graph_AE_use_GAT=False, graph_AE_GAT_dropout=0, graph_AE_learning_rate=1e-2,
                           graph_AE_embedding_size=16, graph_AE_concat_prev_embed=None, graph_AE_normalize_embed=None,
                           graph_AE_neighborhood_factor=10, graph_AE_retain_weights=False, gat_multi_heads=2, gat_hid_embed=64,
                           clustering_louvain_only=False, clustering_embed="graph", clustering_method="KMeans", seed=0,
                           cluster_AE_batch_size=12800, cluster_AE_epoch=4, cluster_AE_learning_rate=1e-3, cluster_AE_regu_strength=0.9,
                           cluster_AE_dropout_prob=0)
    model = mod.ScGNN2(args, device="cuda", seed=0)
    model.fit(X)
    out = model.predict()
    assert out.shape == X.shape and np.isfinite(out).all() and (out >= 0).all()
    labels = np.asarray(model.cluster_labels)
    assert labels.shape == (n, ) and 2 <= len(set(labels.tolist())) <= 31
    # the Louvain half on its own: communities of the returned kNN graph, modularity well above a random split
    lab, nc = mod.generateLouvainCluster(model.edge_list if not isinstance(model.edge_list[0], torch.Tensor) else model.edge_list, n)
    assert 2 <= nc <= n // 5 and len(lab) == n
