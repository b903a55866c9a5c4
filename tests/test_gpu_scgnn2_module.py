"""The reference-facing module API (ScGNN2 / feature_AE_handler / graph_AE_handler) end to end on the GPU,
against the oracle port run with identical initial weights and noise."""
import argparse

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _args(**over):
    # the argparse defaults of examples/single_modality/imputation/scgnn2.py:24-181 that the hot path reads
    d = dict(total_epoch=0, feature_AE_epoch=[3, 2], feature_AE_batch_size=128, feature_AE_learning_rate=1e-3, feature_AE_regu_strength=0.9,
             feature_AE_dropout_prob=0, feature_AE_concat_prev_embed=None, graph_AE_epoch=3, graph_AE_use_GAT=False, graph_AE_GAT_dropout=0,
             graph_AE_learning_rate=1e-2, graph_AE_embedding_size=16, graph_AE_concat_prev_embed=False, graph_AE_normalize_embed=None,
             graph_AE_neighborhood_factor=10, graph_AE_retain_weights=False, gat_multi_heads=2, gat_hid_embed=64)
    d.update(over)
    return argparse.Namespace(**d)


def test_feature_ae_handler_matches_oracle(cuda):
    from dance_b200.modules.scgnn2 import feature_AE_handler
    from oracle import port
    X = port.synthetic_expression(300, 64, density=0.3, seed=3)
    args = _args()
    param = {"device": cuda, "epoch_num": 0, "total_epoch": 0, "n_feature_orig": 64, "seed": 5}
    emb, recon, ckpt = feature_AE_handler(X, None, args, param)
    # oracle with the very same initial weights
    eng = param["_feature_AE_engine"]
    from dance_b200.engine import FeatureAEEngine
    init = FeatureAEEngine(64, device=cuda, seed=5).state_dict()
    ref = port.FeatureAE(64)
    ref.load_state_dict({k: v.cpu() for k, v in init.items()})
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    for _ in range(3):
        _, z_ref, r_ref = port.feature_ae_epoch(ref, opt, torch.from_numpy(X), 128, "LTMG", 0.9)
    assert emb.shape == (300, 128) and recon.shape == (300, 64)
    assert rel_err(emb, z_ref.numpy()) < 1e-4 and rel_err(recon, r_ref.numpy()) < 1e-4
    for k, v in ckpt["model"].items():
        assert rel_err(v.cpu().numpy(), ref.state_dict()[k].numpy()) < 1e-4, k


@pytest.mark.parametrize("use_gat", [False, True])
def test_graph_ae_handler_runs_and_graph_is_exact(cuda, use_gat):
    from dance_b200.modules.scgnn2 import graph_AE_handler
    from oracle import port
    X = port.synthetic_embedding(400, d=128, n_clusters=4, seed=8)
    args = _args(graph_AE_use_GAT=use_gat)
    param = {"device": cuda, "epoch_num": 0, "seed": 1}
    embed, recon, (edge_index, edge_w), adj = graph_AE_handler(X, None, args, param)
    adj_ref, idx_ref = port.feature2adj(X, 10)
    assert np.array_equal(edge_index[:, 1].reshape(400, 10), idx_ref)          # kNN edge list: bit-exact
    assert (adj != adj_ref).nnz == 0                                           # symmetrised adjacency: exact
    assert embed.shape == (400, 16) and recon.shape == (400, 400) and np.isfinite(embed).all()
    assert rel_err(recon, embed @ embed.T) < 1e-5


def test_scgnn2_fit_pre_em(cuda):
    from dance_b200.modules.scgnn2 import ScGNN2
    from oracle import port
    X = port.synthetic_expression(256, 48, density=0.3, seed=1)
    model = ScGNN2(_args(), device="cuda", seed=0)
    model.fit(X)
    out = model.predict()
    assert out.shape == X.shape and np.isfinite(out).all()
    # one EM iteration (the full loop is exercised in tests/test_gpu_em.py and through the example script)
    em = ScGNN2(_args(total_epoch=1, clustering_louvain_only=False, clustering_embed="graph", clustering_method="KMeans", seed=0,
                      cluster_AE_batch_size=12800, cluster_AE_epoch=2, cluster_AE_learning_rate=1e-3, cluster_AE_regu_strength=0.9,
                      cluster_AE_dropout_prob=0), device="cuda", seed=0)
    em.fit(X)
    assert em.predict().shape == X.shape and np.isfinite(em.predict()).all() and len(em.cluster_labels) == 256


def test_graph_ae_handler_locality_order_is_equivalent(cuda):
    """param["cell_order"] = "locality": the Graph-AE epochs run on a relabelled graph (cells grouped by nearest embedding
    centroid, ops.locality_order) and the outputs are un-permuted — same embedding, edge list and adjacency as the plain order."""
    from types import SimpleNamespace
    from dance_b200 import ops
    from dance_b200.modules import scgnn2 as mod
    from oracle import port
    emb = np.abs(port.synthetic_embedding(3000, d=128, n_clusters=6, seed=3)) * 0.05      # Feature-AE-like: non-negative, O(0.1)
    args = SimpleNamespace(graph_AE_use_GAT=False, graph_AE_GAT_dropout=0, graph_AE_concat_prev_embed=None, graph_AE_retain_weights=False,
                           graph_AE_normalize_embed=None, graph_AE_neighborhood_factor=10, graph_AE_embedding_size=16, graph_AE_learning_rate=1e-2,
                           graph_AE_epoch=4, gat_multi_heads=2, gat_hid_embed=64)
    outs = []
    for order in (None, "locality"):
        param = {"device": cuda, "epoch_num": 0, "seed": 0, "precision": "fp32", "cell_order": order, "cell_order_anchors": 16}
        outs.append(mod.graph_AE_handler(emb, None, args, param))
    (z0, _, e0, a0), (z1, _, e1, a1) = outs
    assert np.isfinite(z0).all() and np.linalg.norm(z0 - z1) / np.linalg.norm(z0) < 1e-4
    assert np.array_equal(e0[0], e1[0]) and (a0 != a1).nnz == 0
    perm, inv = ops.locality_order(torch.from_numpy(emb).to(cuda), n_anchors=16)
    assert torch.equal(inv[perm], torch.arange(3000, device=cuda)) and len(torch.unique(perm)) == 3000
