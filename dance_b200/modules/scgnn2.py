"""scGNN 2.0 on the B200-native engines — host-side mirror of the reference module
``dance/modules/single_modality/imputation/scgnn2.py`` for the stages on the hot path.

Same names, argument meaning and return types as the reference:
  * ``ScGNN2(args, device).fit(x) / .predict() / .score(...)``          scgnn2.py:22-121
  * ``feature_AE_handler(X, TRS, args, param, model_state)``            scgnn2.py:275-335
  * ``graph_AE_handler(X_embed, CCC_graph, args, param)``               scgnn2.py:530-600
numpy in / numpy out at this boundary, everything in between stays in HBM.

Differences that are deliberate and documented in DESIGN.md:
  * the N×N decoder logits, the dense label matrix (scgnn2.py:557) and ``recon_graph`` are never
    materialised; ``graph_AE_handler`` returns the ``edgeList`` as an [N·k, 2] index array plus a
    weight array instead of a Python list of tuples, and ``CCC_graph_hat`` as None above
    ``dense_recon_max_cells`` cells;
  * the EM iterations (clustering_handler / cluster_AE_handler, scgnn2.py:138-216, 821-880) are
    SURVEY §8(f) "next" rows and are not built yet: ``fit`` runs the pre-EM stage (the two
    handlers above) and raises for ``total_epoch > 0``.
"""
from __future__ import annotations

import logging
from time import time
from typing import Any, Optional

import numpy as np
import torch

from .. import ops
from ..engine import FeatureAEEngine, GATEngine, GraphAEEngine

logger = logging.getLogger("dance_b200.scgnn2")


def _device(device: str = "auto") -> torch.device:
    if device in ("auto", "cuda", None):
        if not torch.cuda.is_available():
            raise RuntimeError("dance_b200 needs a CUDA device (there is no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"dance_b200 runs on CUDA devices only, got {device!r}")
    return dev


def feature_AE_handler(X, TRS, args, param, model_state=None):
    """Feature autoencoder stage (scgnn2.py:275-335): returns (X_embed, X_recon, checkpoint)."""
    logger.info("Starting Feature AE")
    dev = param["device"]
    batch_size = args.feature_AE_batch_size
    total_epoch = args.feature_AE_epoch[param["epoch_num"] > 0]
    if args.feature_AE_dropout_prob:
        raise NotImplementedError("feature_AE_dropout_prob > 0 (input masking, scgnn2.py:1260) is not built")
    if getattr(args, "feature_AE_concat_prev_embed", None) and param["epoch_num"] > 0:
        raise NotImplementedError("feature_AE_concat_prev_embed is not built")
    Xd = torch.as_tensor(X, dtype=torch.float32).to(dev, non_blocking=True)
    ltmg = None
    if TRS is not None and np.any(TRS):
        ltmg = torch.as_tensor(TRS, dtype=torch.float32).to(dev)
    eng = FeatureAEEngine(Xd.shape[1], device=dev, lr=args.feature_AE_learning_rate, precision=param.get("precision"),
                          seed=param.get("seed"))
    if param["epoch_num"] > 0 and model_state is not None:
        eng.load_state_dict(model_state["model"])
    # regu_type=["LTMG", "noregu"][epoch_num > 0]   (scgnn2.py:314)
    regu = "noregu" if param["epoch_num"] > 0 else "LTMG"
    n = Xd.shape[0]
    z_all = torch.empty(n, eng.EMB, dtype=torch.float32, device=dev)
    r_all = torch.empty(n, Xd.shape[1], dtype=torch.float32, device=dev)
    for epoch in range(total_epoch):
        last = epoch == total_epoch - 1
        loss = eng.train_epoch(Xd, batch_size, regu, args.feature_AE_regu_strength, ltmg, z_all if last else None,
                               r_all if last else None)
        if logger.isEnabledFor(logging.INFO):
            logger.info(f"Epoch: {epoch+1}/{total_epoch}, Average loss: {loss.item() / n:.4f}")
    checkpoint = {"model": eng.state_dict(),
                  "optimizer": {"step": eng.params.step, "exp_avg": eng.params.exp_avg.clone(),
                                "exp_avg_sq": eng.params.exp_avg_sq.clone()}}
    param["_feature_AE_engine"] = eng
    X_embed_out = z_all.cpu().numpy()
    X_recon_out = r_all.cpu().numpy()[:, :param["n_feature_orig"]]
    return X_embed_out, X_recon_out, checkpoint


def build_knn_graph(x_embed: torch.Tensor, neighborhood_factor):
    """feature2adj + preprocess_graph on device (scgnn2.py:650-689, 1191-1198).
    Returns (Â as CSR with the A+I pattern, knn index [N,k] int32, knn fp64 distances)."""
    n = x_embed.shape[0]
    k_tmp = neighborhood_factor if neighborhood_factor > 1 else round(n * neighborhood_factor)
    k = int(k_tmp - 1 if k_tmp == n else k_tmp)
    idx, dist = ops.knn(x_embed, k, include_rank0=False)
    return ops.knn_graph_build(idx), idx, dist


def graph_AE_handler(X_embed, CCC_graph, args, param, dense_recon_max_cells: int = 4096):
    """Graph autoencoder stage, GCN branch (scgnn2.py:530-600): returns (embed, recon_graph, edgeList, adj)."""
    logger.info("Starting Graph AE")
    if args.graph_AE_use_GAT and args.graph_AE_GAT_dropout:
        raise NotImplementedError("graph_AE_GAT_dropout > 0 is not built (the example default is 0)")
    if args.graph_AE_concat_prev_embed and param["epoch_num"] > 0:
        raise NotImplementedError("graph_AE_concat_prev_embed is not built")
    if args.graph_AE_retain_weights:
        raise NotImplementedError("graph_AE_retain_weights permutes node order in the reference (App. B); not built")
    dev = param["device"]
    X = np.asarray(X_embed, dtype=np.float32)
    if args.graph_AE_normalize_embed == "sum1":
        zD = X / np.clip(X.sum(1, keepdims=True), a_min=1, a_max=None)          # scgnn2.py:622-628
    elif args.graph_AE_normalize_embed == "binary":
        zD = (1.0 * (X > np.mean(X, axis=0))).astype(np.float32)
    else:
        zD = X
    xe = torch.from_numpy(X).to(dev)
    A, knn_idx, knn_dist = build_knn_graph(xe, args.graph_AE_neighborhood_factor)
    n = X.shape[0]
    adj_sum = A.nnz - n                                                         # Σ adj_train (no diagonal)
    pos_weight = float(n * n - adj_sum) / adj_sum                               # scgnn2.py:567
    norm = n * n / float((n * n - adj_sum) * 2)                                 # scgnn2.py:568-569
    labels = ops.CSR(A.rowptr, A.colidx, None, A.shape)                         # A + I: pattern of Â, unit entries
    xin = torch.from_numpy(np.ascontiguousarray(zD, dtype=np.float32)).to(dev)
    if args.graph_AE_use_GAT:
        # edge_index = edgeList (i → its k neighbours), directed, no self loops (scgnn2.py:560-563); the kernels
        # index the graph by TARGET node, i.e. the transpose of the regular kNN-list CSR
        k = knn_idx.shape[1]
        src_csr = ops.CSR(torch.arange(0, n * k + 1, k, dtype=torch.int32, device=dev), knn_idx.reshape(-1).contiguous(), None, (n, n))
        T, _ = ops.csr_transpose(src_csr)
        Tt, t_perm = ops.csr_transpose(T)
        geng = GATEngine(X.shape[1], args.gat_hid_embed, args.graph_AE_embedding_size, args.gat_multi_heads, device=dev,
                         lr=args.graph_AE_learning_rate, precision=param.get("precision"), seed=param.get("seed"))
        z = None
        for epoch in range(args.graph_AE_epoch):
            z = geng.train_step(xin, T, Tt, t_perm, labels)                         # loss_function: plain BCE (scgnn2.py:581)
            if logger.isEnabledFor(logging.INFO):
                logger.info(f"Epoch: {epoch+1}/{args.graph_AE_epoch}, Current loss: {geng.loss.item():.4f}")
        param["_graph_AE_engine"] = geng
        return _graph_ae_outputs(z, n, knn_idx, knn_dist, A, dense_recon_max_cells)
    eng = GraphAEEngine(X.shape[1], args.graph_AE_embedding_size, device=dev, lr=args.graph_AE_learning_rate,
                        precision=param.get("precision"), seed=param.get("seed"))
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(param.get("seed") or 0))
    eps = torch.empty(n, args.graph_AE_embedding_size, dtype=torch.float32, device=dev)
    z = None
    for epoch in range(args.graph_AE_epoch):
        eps.normal_(generator=gen)                                              # torch.randn_like(std), scgnn2.py:397
        z, _, _ = eng.train_step(xin, A, labels, norm, pos_weight, eps)
        if logger.isEnabledFor(logging.INFO):
            logger.info(f"Epoch: {epoch+1}/{args.graph_AE_epoch}, Current loss: {eng.loss.item():.4f}")
    param["_graph_AE_engine"] = eng
    return _graph_ae_outputs(z, n, knn_idx, knn_dist, A, dense_recon_max_cells)


def _graph_ae_outputs(z, n, knn_idx, knn_dist, A, dense_recon_max_cells):
    embed_out = z.cpu().numpy()
    recon = (z @ z.t()).cpu().numpy() if n <= dense_recon_max_cells else None   # InnerProductDecoder output, small N only
    k = knn_idx.shape[1]
    edge_index = np.stack([np.repeat(np.arange(n), k), knn_idx.cpu().numpy().reshape(-1).astype(np.int64)], 1)
    edge_w = 1.0 / (knn_dist.cpu().numpy().reshape(-1) + 1e-16)                 # scgnn2.py:686
    adj = A.to_scipy()
    adj.data[:] = 1.0
    adj.setdiag(0)
    adj.eliminate_zeros()
    return embed_out, recon, (edge_index, edge_w), adj


class ScGNN2:
    """Drop-in for ``dance.modules.single_modality.imputation.scgnn2.ScGNN2`` (pre-EM stage)."""

    def __init__(self, args, device: str = "auto", precision: Optional[str] = None, seed: Optional[int] = None):
        self.args = args
        self.device = _device(device)
        self.precision = precision
        self.seed = seed

    def fit(self, x: np.ndarray):
        args = self.args
        epochs = args.total_epoch
        param = {"device": self.device, "tik": time(), "precision": self.precision, "seed": self.seed}
        logger.info(f"Using device: {param['device']}")
        trs_mat = None  # the reference passes an all-zero TRS (scgnn2.py:40)
        logger.info("Pre EM runs")
        param["epoch_num"] = 0
        param["total_epoch"] = epochs
        param["n_feature_orig"] = x.shape[1]
        param["x_dropout"] = x
        x_embed, x_feature_recon, model_state = feature_AE_handler(x, trs_mat, args, param)
        graph_embed, _, edge_list, adj = graph_AE_handler(x_embed, None, args, param)
        self.x_embed, self.graph_embed, self.edge_list, self.adj = x_embed, graph_embed, edge_list, adj
        self.model_state = model_state
        if epochs > 0:
            raise NotImplementedError(
                "EM iterations (clustering_handler, graph_celltype_regu_handler, cluster_AE_handler; reference "
                "scgnn2.py:56-66) are SURVEY §8(f) 'next' rows and are not built yet; run with total_epoch=0")
        self.x_imputed = x_feature_recon

    def predict(self, x: Optional[Any] = None) -> np.ndarray:
        return self.x_imputed

    def score(self, true_expr, imputed_expr, mask=None, metric="MSE", log1p=True, test_idx=None):
        """Same scoring as the reference (scgnn2.py:73-121): 'RMSE' | 'PCC' | 'MRE'."""
        allowd_metrics = {"RMSE", "PCC", "MRE"}
        if metric not in allowd_metrics:
            raise ValueError("scoring metric %r." % allowd_metrics)
        if test_idx is None:
            test_idx = range(len(true_expr))
        true_target = true_expr[test_idx].to(self.device)
        imputed_target = imputed_expr[test_idx].to(self.device)
        if log1p:
            imputed_target = torch.log1p(imputed_target)
        if mask is not None:
            imputed_target[mask[test_idx]] = true_target[mask[test_idx]].to(imputed_target.dtype)
        if metric == "RMSE":
            return np.sqrt(torch.nn.functional.mse_loss(true_target, imputed_target).item())
        elif metric == "PCC":
            return np.corrcoef(true_target.cpu()[~mask[test_idx]], imputed_target.cpu()[~mask[test_idx]])[0, 1]
        elif metric == "MRE":
            actual = true_target.cpu()[~mask[test_idx]]
            predicted = imputed_target.cpu()[~mask[test_idx]]
            abs_error = torch.abs(predicted - actual)
            abs_actual = torch.abs(actual)
            abs_actual[abs_actual < 1e-10] = 1e-10
            return torch.mean(abs_error / abs_actual).item()
