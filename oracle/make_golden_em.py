"""Golden fixture for the scGNN EM-iteration stages, produced by the REFERENCE's own functions executed
through ``oracle.ref_loader`` — TEST INFRASTRUCTURE, run in the build container:  ``python -m oracle.make_golden_em``.

Covers  graph_celltype_regu_handler / normalize_cell_cell_matrix / generateCelltypeRegu (scgnn2.py:716-752),
loss_function_graph(regularizer_type="Celltype") value and gradient (scgnn2.py:1316-1326), train_handler's L1 term
(scgnn2.py:1268-1274) and cluster_AE_handler end to end (scgnn2.py:821-880), trimClustering / cluster_output_handler
(scgnn2.py:218-254).  generateLouvainCluster needs igraph (absent): only its networkx half is exercised, see tests.
"""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import torch

from . import port, ref_loader

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    ref = ref_loader.scgnn2()
    rng = np.random.default_rng(21)
    n, g = 240, 48
    x_dropout = port.synthetic_expression(n, g, density=0.3, seed=9)                  # the matrix fit() was called with
    X_recon = np.maximum(x_dropout + rng.normal(size=(n, g)).astype(np.float32) * 0.2, 0).astype(np.float32)   # Feature-AE output stand-in
    emb = port.synthetic_embedding(n, d=16, n_clusters=3, seed=4)
    _, adj_train, edge_list = ref.feature2adj(emb, 6, False)
    adj_train = sp.csr_matrix(adj_train)
    adj_train.sort_indices()
    labels = (np.arange(n) * 3 // n).astype(np.int64)                                 # three clusters of 80 cells
    rng.shuffle(labels)
    _, lists = ref.cluster_output_handler(labels.tolist())
    adjdense, celltypesample = ref.graph_celltype_regu_handler(adj_train, labels.tolist())
    adjdense, celltypesample = np.asarray(adjdense, dtype=np.float64), np.asarray(celltypesample, dtype=np.float64)

    # per-cluster column sums of the two dense regularisers — what the Cluster-AE loss takes from them
    w_graph, w_ct = np.zeros(n), np.zeros(n)
    for members in lists:
        w_graph[members] = adjdense[np.ix_(members, members)].sum(0)
        w_ct[members] = celltypesample[np.ix_(members, members)].sum(0)

    torch.manual_seed(5)
    fae = ref.Feature_AE(dim=g)
    state = {k: v.detach().clone() for k, v in fae.state_dict().items()}
    out = dict(x_dropout=x_dropout, X_recon=X_recon, adj_indptr=adj_train.indptr, adj_indices=adj_train.indices, labels=labels,
               w_graph=w_graph, w_celltype=w_ct, **{"sd_" + k: v.numpy() for k, v in state.items()})

    # loss_function_graph("Celltype") on cluster 0 as ONE batch: value and gradient w.r.t. the reconstruction, plus the L1 term
    members = lists[0]
    param = {"device": "cpu", "epoch_num": 1, "total_epoch": 2, "n_feature_orig": g, "dataloader_kwargs": {}}
    model = ref.Cluster_AE(dim=g)
    model.load_state_dict(state)
    data = torch.from_numpy(X_recon[members])
    z, recon = model(data)
    recon.retain_grad()
    regu = {"graph_regu": torch.from_numpy(adjdense[np.ix_(members, members)]).float(),
            "celltype_regu": torch.from_numpy(celltypesample[np.ix_(members, members)]).float(),
            "x_dropout": torch.from_numpy(x_dropout[members]), "LTMG_regu": torch.zeros(len(members), g)}
    loss = ref.loss_function_graph(recon, data.view(-1, g), regulationMatrix=regu, regu_strength=0.9, regularizer_type="Celltype",
                                   param=param)
    l1 = sum(p.abs().sum() for p in model.parameters())
    (loss + l1).backward()
    out.update(batch_members=np.asarray(members), batch_recon=recon.detach().numpy(), batch_loss=float(loss.item()), batch_l1=float(l1.item()),
               batch_grad_recon=recon.grad.numpy(),
               **{"batch_grad_" + k: p.grad.numpy() for k, p in model.named_parameters()})

    # cluster_AE_handler end to end (3 epochs per cluster, Adam lr 1e-3)
    args = SimpleNamespace(cluster_AE_batch_size=12800, cluster_AE_epoch=3, cluster_AE_learning_rate=1e-3, cluster_AE_regu_strength=0.9,
                           cluster_AE_dropout_prob=0)
    param["impute_regu"] = (adjdense, celltypesample)
    param["x_dropout"] = x_dropout
    recon_out = ref.cluster_AE_handler(X_recon, np.zeros_like(X_recon), lists, args, param, {"model": state})
    out.update(cluster_recon=recon_out, cluster_epochs=3)

    # trimClustering on a labelling with tiny clusters (labels 0..k-1 all present, as the reference requires)
    lab2 = np.concatenate([np.full(40, 0), np.full(3, 1), np.full(30, 2), np.full(6, 3), np.full(5, 4), np.full(1, 5)])
    out.update(trim_in=lab2, trim_out=np.asarray(ref.trimClustering(lab2.tolist(), minMemberinCluster=5, maxClusterNumber=30)))
    np.savez_compressed(OUT / "scgnn_em.npz", **out)
    print("wrote", OUT / "scgnn_em.npz", {k: np.asarray(v).shape for k, v in out.items() if k.startswith(("cluster", "batch_loss", "w_"))})


if __name__ == "__main__":
    main()
