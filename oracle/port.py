"""CPU restatement ("port") of the reference's algorithm for the scGNN message-passing
hot path.  TEST INFRASTRUCTURE: imported only by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Each function cites the reference file:line it follows (paths relative to the
reference repository root).  Parity status: PINNED — every function here is checked
against the reference's own code executed through ``oracle.ref_loader`` (in the build
container, ``tests/test_oracle_vs_reference.py``) and against the committed fixtures in
``tests/golden/`` generated from that reference (``oracle/make_golden.py``); the
normalize / distance functions are additionally pinned by the reference's own
known-answer tests (tests/utils/test_matrix.py, tests/transforms/test_normalize.py).

Implementation language: numpy / scipy / torch-CPU, i.e. the same libraries the
reference's CPU path runs on, so timing it is timing the reference's CPU arithmetic.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F
from scipy.spatial import distance


# --------------------------------------------------------------------------- kNN graph
def knn_indices(X: np.ndarray, k: int, block: int = 1024, return_dist: bool = False):
    """Sorted ranks 1..k of every row under fp64 euclidean distance.

    Follows calculateKNNgraphDistanceMatrixStatsSingleThread (scgnn2.py:675-689): per row
    ``distance.cdist(row, X, "euclidean")`` (fp64 on the fp32 features), ``argsort``, take
    ``res[0][1..k]``.  Rows are processed in blocks (identical per-pair arithmetic) and the
    argsort is made tie-stable (``kind="stable"`` → ties by smaller index; the reference's
    quicksort leaves tie order unspecified).
    """
    n = X.shape[0]
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64) if return_dist else None
    for i0 in range(0, n, block):
        d = distance.cdist(X[i0:i0 + block], X, "euclidean")
        order = np.argsort(d, axis=1, kind="stable")[:, 1:k + 1]
        idx[i0:i0 + block] = order
        if return_dist:
            dist[i0:i0 + block] = np.take_along_axis(d, order, axis=1)
    return (idx, dist) if return_dist else idx


def feature2adj(X_embed: np.ndarray, neighborhood_factor, block: int = 1024):
    """feature2adj with retain_weights=False (scgnn2.py:650-672).

    Returns (adj_train CSR 0/1 symmetric without diagonal, knn index array).  The reference
    goes through networkx ``from_dict_of_lists`` (undirected ⇒ union-symmetrised, rows in
    natural order, App. B) and then clears the diagonal; the same matrix is built here with
    scipy.
    """
    n = X_embed.shape[0]
    k_tmp = neighborhood_factor if neighborhood_factor > 1 else round(n * neighborhood_factor)
    k = int(k_tmp - 1 if k_tmp == n else k_tmp)
    idx = knn_indices(X_embed, k, block=block)
    rows = np.repeat(np.arange(n, dtype=np.int64), k)
    cols = idx.reshape(-1)
    a = sp.csr_matrix((np.ones(rows.size, dtype=np.float64), (rows, cols)), shape=(n, n))
    a = a + a.T
    a.data[:] = 1.0
    a.setdiag(0)
    a.eliminate_zeros()
    a.sort_indices()
    return a.tocsr(), idx


def preprocess_graph(adj_train: sp.spmatrix) -> sp.csr_matrix:
    """Â = D^-1/2 (A + I) D^-1/2 as fp32 CSR with sorted columns (scgnn2.py:1191-1198,1205)."""
    adj = sp.coo_matrix(adj_train)
    adj_ = adj + sp.eye(adj.shape[0])
    rowsum = np.array(adj_.sum(1))
    d_inv_sqrt = sp.diags(np.power(rowsum, -0.5).flatten())
    out = adj_.dot(d_inv_sqrt).transpose().dot(d_inv_sqrt).tocsr().astype(np.float32)
    out.sort_indices()
    return out


def gae_norm_constants(adj_train: sp.spmatrix) -> Tuple[float, float]:
    """pos_weight and norm of graph_AE_handler (scgnn2.py:567-569)."""
    n = adj_train.shape[0]
    s = adj_train.sum()
    pos_weight = float(n * n - s) / s
    norm = n * n / float((n * n - s) * 2)
    return pos_weight, norm


def to_torch_sparse(m: sp.spmatrix) -> torch.Tensor:
    """sparse_mx_to_torch_sparse_tensor (scgnn2.py:1201-1209)."""
    m = m.tocoo().astype(np.float32)
    idx = torch.from_numpy(np.vstack((m.row, m.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data), torch.Size(m.shape)).coalesce()


# --------------------------------------------------------------------------- Graph AE (GCN branch)
def graph_conv(x: torch.Tensor, weight: torch.Tensor, adj: torch.Tensor, act=None) -> torch.Tensor:
    """GraphConvolution.forward with dropout 0 (scgnn2.py:497-502)."""
    support = torch.mm(x, weight)
    out = torch.spmm(adj, support)
    return act(out) if act is not None else out


def gae_loss(preds, labels, mu, logvar, n_nodes, norm, pos_weight):
    """gae_loss_function (scgnn2.py:603-615)."""
    cost = norm * F.binary_cross_entropy_with_logits(preds, labels, pos_weight=labels * pos_weight)
    if logvar is None:
        return cost
    kld = -0.5 / n_nodes * torch.mean(torch.sum(1 + 2 * logvar - mu.pow(2) - logvar.exp().pow(2), 1))
    return cost + kld


def graph_ae_gcn_forward(x, w1, w2, w3, adj, eps: Optional[torch.Tensor]):
    """Graph_AE.forward with use_GAT=False (scgnn2.py:389-412): returns (z, mu, logvar, hidden1).

    ``eps`` is the reparameterisation noise (``torch.randn_like(std)``, scgnn2.py:397); pass
    None for eval mode (z = mu).
    """
    hidden1 = graph_conv(x, w1, adj, F.relu)
    mu = graph_conv(hidden1, w2, adj)
    logvar = graph_conv(hidden1, w3, adj)
    z = mu if eps is None else eps.mul(torch.exp(logvar)).add(mu)
    return z, mu, logvar, hidden1


def graph_ae_gcn_loss(x, w1, w2, w3, adj_norm_sp: sp.spmatrix, adj_train: sp.spmatrix, eps=None, dense_labels=True):
    """One forward of the GCN branch of graph_AE_handler (scgnn2.py:555-590): loss + tensors.

    The label matrix is the dense (A + I) of scgnn2.py:557 — only feasible for small n.
    """
    n = x.shape[0]
    adj = to_torch_sparse(adj_norm_sp)
    pos_weight, norm = gae_norm_constants(adj_train)
    z, mu, logvar, hidden1 = graph_ae_gcn_forward(x, w1, w2, w3, adj, eps)
    labels = torch.from_numpy((adj_train + sp.eye(n)).toarray()).float()
    preds = torch.mm(z, z.t())  # InnerProductDecoder, identity activation (scgnn2.py:423-426)
    loss = gae_loss(preds, labels, mu, logvar, n, norm, pos_weight)
    return loss, z, mu, logvar, hidden1


# --------------------------------------------------------------------------- Graph AE (GAT branch)
def gat_layer(x, edge_index, proj_w, skip_w, a_src, a_trg, bias, concat: bool, act=None):
    """GATLayer.forward with dropout 0 (scgnn2.py:989-1051 and helpers :1057-1215).

    edge_index[0] = source nodes, edge_index[1] = target nodes; scores use LeakyReLU(0.2), the softmax shift is
    the GLOBAL max over all edges and heads (:1076), the denominator gets +1e-16 (:1086), the skip connection is
    the projected one (the reference adds the raw input only when FIN == FOUT, :1194-1197)."""
    nh, F_ = a_src.shape[1], a_src.shape[2]
    n = x.shape[0]
    src, trg = edge_index[0], edge_index[1]
    proj = (x @ proj_w.t()).view(-1, nh, F_)
    s_src = (proj * a_src).sum(-1)
    s_trg = (proj * a_trg).sum(-1)
    scores = torch.nn.functional.leaky_relu(s_src.index_select(0, src) + s_trg.index_select(0, trg), 0.2)
    ex = (scores - scores.max()).exp()
    denom = torch.zeros(n, nh, dtype=ex.dtype).index_add_(0, trg, ex)
    att = (ex / (denom.index_select(0, trg) + 1e-16)).unsqueeze(-1)
    out = torch.zeros(n, nh, F_, dtype=x.dtype).index_add_(0, trg, proj.index_select(0, src) * att)
    if out.shape[-1] == x.shape[-1]:
        out = out + x.unsqueeze(1)
    else:
        out = out + (x @ skip_w.t()).view(-1, nh, F_)
    out = out.view(-1, nh * F_) if concat else out.mean(dim=1)
    if bias is not None:
        out = out + bias
    return act(out) if act is not None else out


def graph_ae_gat_forward(x, edge_index, sd):
    """Graph_AE.encode_gat (scgnn2.py:385-386): 2 GAT layers — concat+ELU, then head-mean — from a reference
    state_dict (keys gat.gat_net.{l}.*)."""
    h = x
    for l, (concat, act) in enumerate(((True, torch.nn.functional.elu), (False, None))):
        pre = f"gat.gat_net.{l}."
        h = gat_layer(h, edge_index, sd[pre + "linear_proj.weight"], sd[pre + "skip_proj.weight"], sd[pre + "scoring_fn_source"],
                      sd[pre + "scoring_fn_target"], sd[pre + "bias"], concat, act)
    return h


# --------------------------------------------------------------------------- Feature AE
class FeatureAE(torch.nn.Module):
    """Feature_AE (scgnn2.py:338-370): dim→512→128→512→dim, ReLU after every layer."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        self.fc1 = torch.nn.Linear(dim, 512)
        self.fc2 = torch.nn.Linear(512, 128)
        self.fc3 = torch.nn.Linear(128, 512)
        self.fc4 = torch.nn.Linear(512, dim)

    def forward(self, x):
        z = F.relu(self.fc2(F.relu(self.fc1(x.view(-1, self.dim)))))
        return z, F.relu(self.fc4(F.relu(self.fc3(z))))


def feature_ae_loss(recon, x, regularizer_type="noregu", regu_strength=0.9, ltmg=None):
    """loss_function_graph, branches 'noregu' and 'LTMG' (scgnn2.py:1298-1315)."""
    bce = F.mse_loss(recon, x, reduction="sum")
    if regularizer_type == "noregu":
        return bce
    if regularizer_type == "LTMG":
        return (1 - regu_strength) * bce + regu_strength * (F.mse_loss(recon, x, reduction="none") * ltmg).sum()
    raise ValueError(regularizer_type)


def feature_ae_epoch(model: FeatureAE, optimizer, X: torch.Tensor, batch_size: int, regularizer_type="LTMG",
                     regu_strength=0.9, ltmg: Optional[torch.Tensor] = None):
    """One epoch of train_handler with masked_prob = 0 (scgnn2.py:1254-1293); returns
    (sum of batch losses, z_all, recon_all)."""
    model.train()
    total = 0.0
    zs, rs = [], []
    n = X.shape[0]
    for b0 in range(0, n, batch_size):
        data = X[b0:b0 + batch_size]
        t = ltmg[b0:b0 + batch_size] if ltmg is not None else torch.zeros_like(data)
        optimizer.zero_grad()
        z, recon = model(data)
        loss = feature_ae_loss(recon, data, regularizer_type, regu_strength, t)
        loss.backward()
        total += loss.item()
        optimizer.step()
        zs.append(z.detach())
        rs.append(recon.detach())
    return total, torch.cat(zs, 0), torch.cat(rs, 0)


# --------------------------------------------------------------------------- scDeepSort path (DGL-backed in the reference)
# Parity status of this block: UNPINNED — the reference routes these through DGL 1.1.3 (absent here), so
# they are restated from the reference text (cell_feature_graph.py:34-79, gnn.py:62-96, scdeepsort.py:142-250) and the
# DGL semantics of SURVEY App. A; no reference execution or golden vector exists for them.
def cell_feature_graph(feat: np.ndarray, normalize_edges: bool = True):
    """CellFeatureGraph.__call__ edge list (cell_feature_graph.py:38-69): (src, dst, w[E,1]) with gene nodes first."""
    feat = np.asarray(feat, dtype=np.float32)
    n, g = feat.shape
    row, col = np.nonzero(feat)
    edata = feat[row, col].ravel()[:, None]
    row = row + g
    col, row = np.hstack((col, row)), np.hstack((row, col))
    w = torch.from_numpy(np.vstack((edata, edata)).astype(np.float32))
    src, dst = torch.from_numpy(row.astype(np.int64)), torch.from_numpy(col.astype(np.int64))
    if normalize_edges:
        in_deg = torch.bincount(dst, minlength=n + g)
        order = torch.argsort(dst, stable=True)
        bounds = torch.zeros(n + g + 1, dtype=torch.int64)
        bounds[1:] = torch.cumsum(in_deg, 0)
        for i in range(n + g):                      # the reference's per-node loop (:64-68), in fp32 like torch
            eidx = order[bounds[i]:bounds[i + 1]]
            if eidx.numel() > 0:
                ew = w[eidx]
                w[eidx] = in_deg[i] * ew / ew.sum()
    nodes = torch.arange(n + g, dtype=torch.int64)
    return torch.cat([src, nodes]), torch.cat([dst, nodes]), torch.cat([w, torch.ones(n + g, 1)])


def adaptive_sage_neighbour_mean(src, dst, w, h, alpha, n_genes: int):
    """update_all(message_func, fn.mean) of AdaptiveSAGE (gnn.py:62-90) on the full graph: the value the reference
    stores in "neigh" (and never uses)."""
    src_gene, dst_gene = src < n_genes, dst < n_genes
    idx = torch.full_like(src, n_genes + 1)
    idx = torch.where(src_gene & ~dst_gene, src, idx)
    idx = torch.where(dst_gene & ~src_gene, dst, idx)
    idx = torch.where(dst_gene & src_gene, torch.full_like(src, n_genes), idx)
    m = h[src] * alpha[idx] * w
    out = torch.zeros_like(h).index_add_(0, dst, m)
    deg = torch.bincount(dst, minlength=h.shape[0]).clamp(min=1).unsqueeze(1)
    return out / deg


class ScDeepSortNet(torch.nn.Module):
    """GNN (scdeepsort.py:26-88) with one AdaptiveSAGE layer as it actually computes (gnn.py:84-96): the output
    depends on the destination features only."""

    def __init__(self, dim_in, dim_hid, n_labels, gene_num):
        super().__init__()
        self.alpha = torch.nn.Parameter(torch.ones(gene_num + 2, 1))
        self.sage_linear = torch.nn.Linear(dim_in, dim_hid)
        torch.nn.init.xavier_uniform_(self.sage_linear.weight, gain=torch.nn.init.calculate_gain("relu"))
        self.linear = torch.nn.Linear(dim_hid, n_labels)
        torch.nn.init.xavier_uniform_(self.linear.weight, gain=torch.nn.init.calculate_gain("relu"))

    def forward(self, h_dst):
        return self.linear(torch.relu(self.sage_linear(h_dst)))


def scdeepsort_epoch(net, optimizer, feats, labels, batch_index_lists):
    """cal_loss (scdeepsort.py:213-250) for explicit batches: CrossEntropyLoss(reduction='sum'), Adam."""
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")
    total_loss = total_size = 0.0
    for idx in batch_index_lists:
        idx = torch.as_tensor(idx, dtype=torch.int64)
        loss = loss_fn(net(feats[idx]), labels[idx])
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        total_size += idx.numel()
        total_loss += loss.item() * idx.numel()
    return total_loss / total_size


# --------------------------------------------------------------------------- preprocessing
def normalize_total(X: np.ndarray, target_sum: Optional[float] = None, exclude_highly_expressed: bool = False,
                    max_fraction: float = 0.05) -> np.ndarray:
    """scanpy.pp.normalize_total 1.10.1 on a dense matrix (called at
    transforms/normalize.py:618-620 and through AnnDataTransform, interface.py:67-68).
    Pinned by the reference's tests/transforms/test_normalize.py:8-30."""
    X = np.array(X, dtype=np.float32, copy=True)
    counts = X.sum(1)
    if exclude_highly_expressed:
        hi = (X > counts[:, None] * max_fraction).sum(0) > 0
        counts = X[:, ~hi].sum(1)
    if target_sum is None:
        target_sum = np.median(counts[counts > 0], axis=0)
    scale = counts / target_sum
    scale = scale + (scale == 0)  # zero-count cells are left unchanged (scanpy >= 1.10.1)
    return (X / scale[:, None]).astype(np.float32)


def log1p(X: np.ndarray, base: Optional[float] = None) -> np.ndarray:
    """scanpy.pp.log1p (called at transforms/normalize.py:563); pinned by test_normalize.py:33-43."""
    out = np.log1p(np.asarray(X, dtype=np.float32))
    if base is not None:
        out = out / np.float32(np.log(base))
    return out


def pairwise_euclidean(X: np.ndarray) -> np.ndarray:
    """dance.utils.matrix.pairwise_distance(x, 0) (utils/matrix.py:100-105,164-180): per pair
    (a-b)² in fp32, accumulated in fp64 (numba unifies ``sum = 0`` with the fp32 terms to
    float64), sqrt, cast to fp32."""
    X = np.asarray(X, dtype=np.float32)
    diff = X[:, None, :] - X[None, :, :]
    sq = (diff * diff).astype(np.float32)
    s = np.zeros(sq.shape[:2], dtype=np.float64)
    for c in range(X.shape[1]):  # sequential accumulation order of the numba loop
        s += sq[:, :, c].astype(np.float64)
    return np.sqrt(s).astype(np.float32)


def matrix_normalize(mat: np.ndarray, mode: str = "normalize", axis: int = 0, eps: float = -1.0) -> np.ndarray:
    """dance.utils.matrix.normalize (utils/matrix.py:8-67); pinned by tests/utils/test_matrix.py:9-29."""
    opts = {"axis": axis, "keepdims": True}
    shift = 0
    if mode == "standardize":
        shift = -mat.mean(**opts)
    elif mode == "minmax":
        min_vals = mat.min(**opts)
        shift = -min_vals
    if mode == "normalize":
        denom = mat.sum(**opts)
    elif mode == "standardize":
        denom = mat.std(**opts)
    elif mode == "minmax":
        denom = mat.max(**opts) - min_vals
    elif mode == "l2":
        denom = (mat**2).sum(**opts)**0.5
    else:
        denom = None
    if denom is None:
        denom = 1
    elif eps == -1:
        denom = np.where(denom == 0, 1, denom)
    elif eps > 0:
        denom = denom + eps
    else:
        raise ValueError(f"Invalid {eps=!r}. Must be positive or -1, the later set zero entries to one.")
    return (mat + shift) / denom


# --------------------------------------------------------------------------- synthetic data
# ---------------------------------------------------------------------------------------------------------------
# SpaGCN (reference dance/modules/spatial/spatial_domain/spagcn.py)
# ---------------------------------------------------------------------------------------------------------------
def spagcn_calculate_p(adj: np.ndarray, l: float) -> float:
    """spagcn.py:249-251 — mean row sum of exp(-adj²/2l²), minus the self term."""
    adj_exp = np.exp(-1 * (adj**2) / (2 * (l**2)))
    return float(np.mean(np.sum(adj_exp, 1)) - 1)


def spagcn_search_l(p, adj, start=0.01, end=1000, tol=0.01, max_run=100):
    """spagcn.py:254-287 — bisection on l."""
    run = 0
    p_low, p_high = spagcn_calculate_p(adj, start), spagcn_calculate_p(adj, end)
    if p_low > p + tol or p_high < p - tol:
        return None
    if abs(p_low - p) <= tol:
        return start
    if abs(p_high - p) <= tol:
        return end
    while (p_low + tol) < p < (p_high - tol):
        run += 1
        if run > max_run:
            return None
        mid = (start + end) / 2
        p_mid = spagcn_calculate_p(adj, mid)
        if abs(p_mid - p) <= tol:
            return mid
        if p_mid <= p:
            start, p_low = mid, p_mid
        else:
            end, p_high = mid, p_mid
    return None


def spagcn_forward(X, adj, W, b, mu, alpha: float = 0.2):
    """GraphConvolution.forward :357-363 + SimpleGCDEC.forward :391-397 (torch tensors, autograd-capable)."""
    z = torch.mm(adj, torch.mm(X, W)) + b
    q = 1.0 / ((1.0 + torch.sum((z.unsqueeze(1) - mu)**2, dim=2) / alpha) + 1e-8)
    q = q**(alpha + 1.0) / 2.0
    q = q / torch.sum(q, dim=1, keepdim=True)
    return z, q


def spagcn_target(q):
    """target_distribution :408-425."""
    p = q**2 / torch.sum(q, dim=0)
    return p / torch.sum(p, dim=1, keepdim=True)


def spagcn_kl(p, q):
    """loss_function :399-406."""
    return torch.mean(torch.sum(p * torch.log(p / (q + 1e-6)), dim=1))


def spagcn_group_means(features: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Cluster centres in sorted-label order (groupby("Group").mean(), :499-503)."""
    return np.stack([features[y == c].mean(0) for c in np.unique(y)]).astype(np.float32)


def spagcn_fit(X, adj, W, b, init_y, lr, epochs, update_interval=3, weight_decay=0.0, opt="admin", tol=1e-3, alpha=0.2,
               train_mu=False, mu=None):
    """The training loop of SimpleGCDEC.fit (:505-534; ``train_mu=False`` — mu is not in the optimiser) or of
    fit_with_init (:563-574; ``train_mu=True``, no stopping rule).  Returns (W, b, mu, epochs_run)."""
    X, adj = torch.as_tensor(X), torch.as_tensor(adj)
    W = torch.nn.Parameter(torch.as_tensor(W).clone())
    b = torch.nn.Parameter(torch.as_tensor(b).clone())
    with torch.no_grad():
        feats = (torch.mm(adj, torch.mm(X, W)) + b).numpy()
    mu_t = torch.nn.Parameter(torch.as_tensor(spagcn_group_means(feats, np.asarray(init_y))))
    params = [W, b] + ([mu_t] if train_mu else [])
    optim = torch.optim.SGD(params, lr=lr, momentum=0.9) if opt == "sgd" else torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)
    y_last = np.asarray(init_y)
    done = 0
    for epoch in range(epochs):
        if epoch % update_interval == 0:
            with torch.no_grad():
                p = spagcn_target(spagcn_forward(X, adj, W, b, mu_t, alpha)[1])
        optim.zero_grad()
        _, q = spagcn_forward(X, adj, W, b, mu_t, alpha)
        spagcn_kl(p, q).backward()
        optim.step()
        done = epoch + 1
        if not train_mu:
            y = torch.argmax(q, dim=1).numpy()
            delta = np.sum(y != y_last).astype(np.float32) / X.shape[0]
            y_last = y
            if epoch > 0 and (epoch - 1) % update_interval == 0 and delta < tol:
                break
    return W.detach().numpy(), b.detach().numpy(), mu_t.detach().numpy(), done


def feature_feature_graph(feat: np.ndarray, threshold: float = 0.3, positive_only: bool = False, normalize_edges: bool = True,
                          score_func: str = "pearson", score_func_kwargs=None):
    """FeatureFeatureGraph (feature_feature_graph.py:45-87) with the pearson / spearman / rbf score; dgl.graph + EdgeWeightNorm("both")
    restated (dgl 1.1.3, un-vendored): weight_e = outdeg_w(src)^-0.5 · indeg_w(dst)^-0.5 · w_e on unit weights.
    Returns (src int32, dst int32, w fp32, adj fp32 after thresholding)."""
    if score_func == "pearson":
        adj = np.corrcoef(feat.T)
    elif score_func == "spearman":
        from scipy.stats import spearmanr
        adj = spearmanr(feat, axis=0)[0]
    elif score_func == "rbf":
        norm_vec = np.power(feat, 2).sum(0, keepdims=True)
        dist_mat = np.sqrt((norm_vec + norm_vec.T - 2 * feat.T @ feat).clip(0))
        kw = dict(score_func_kwargs or {})
        mode, scale = kw.get("scale_mode", "med_dist"), kw.get("denom_scale", 1.0)          # dist_to_rbf, utils/matrix.py:70-97
        denom = {"med_dist": lambda: np.median(dist_mat) * scale, "ind_med_dist": lambda: np.median(dist_mat, axis=1, keepdims=True) * scale,
                 "scale": lambda: scale}[mode]()
        adj = np.exp(-dist_mat / denom)
    else:
        raise ValueError(score_func)
    adj = adj.astype(np.float32)
    adj[np.logical_and(adj > -threshold, adj < threshold)] = 0
    if positive_only:
        adj[adj < 0] = 0
    coo = sp.coo_matrix(adj)
    src, dst = coo.row.astype(np.int32), coo.col.astype(np.int32)
    w = torch.ones(len(src), dtype=torch.float32)
    if normalize_edges:
        g = adj.shape[0]
        out_deg = torch.zeros(g).index_add_(0, torch.from_numpy(src).long(), w)
        in_deg = torch.zeros(g).index_add_(0, torch.from_numpy(dst).long(), w)
        w = torch.pow(out_deg, -0.5)[torch.from_numpy(src).long()] * torch.pow(in_deg, -0.5)[torch.from_numpy(dst).long()] * w
    return src, dst, w.numpy(), adj


def weighted_graphconv(x, src, dst, w_e, W, b, norm: str = "both", agg: str = "sum", act=None):
    """graph-sc's in-tree WeightedGraphConv.forward (modules/single_modality/clustering/graphsc.py:428-484): out-degree^-1/2 on
    the sources (norm="both"), W first, messages h_src·w_e, sum | mean over in-edges, in-degree scaling, bias, activation.
    Degrees are structural (edge counts, clamped to 1).  torch tensors; src/dst int64, w_e [E] or [E,1]."""
    n = x.shape[0]
    indeg = torch.bincount(dst, minlength=n).float().clamp(min=1)
    outdeg = torch.bincount(src, minlength=n).float().clamp(min=1)
    h = x
    if norm == "both":
        h = h * outdeg.pow(-0.5)[:, None]
    h = h @ W
    m = h[src] * w_e.reshape(-1, 1)
    rst = torch.zeros(n, W.shape[1], dtype=h.dtype).index_add(0, dst, m)
    if agg == "mean":
        rst = rst / indeg[:, None]
    if norm != "none":
        rst = rst * (indeg.pow(-0.5) if norm == "both" else 1.0 / indeg)[:, None]
    rst = rst + b
    return act(rst) if act is not None else rst


def umap_connectivities(knn_idx: np.ndarray, knn_dist: np.ndarray) -> sp.csr_matrix:
    """scanpy 1.10.1 ``_connectivity.umap`` → umap-learn 0.5 ``fuzzy_simplicial_set(set_op_mix_ratio=1,
    local_connectivity=1)`` restated loop by loop (smooth_knn_dist, compute_membership_strengths, fuzzy union);
    both packages are un-vendored third-party dependencies of the reference (requirements.txt:19) — parity unpinned."""
    n, k = knn_idx.shape
    dist = knn_dist.astype(np.float32)
    target = np.log2(k)
    rho = np.zeros(n, np.float32)
    sig = np.zeros(n, np.float32)
    mean_all = np.float32(dist.mean())
    for i in range(n):
        lo, hi, mid = 0.0, np.inf, 1.0
        nz = dist[i][dist[i] > 0.0]
        if nz.shape[0] >= 1:
            rho[i] = nz[0]
        elif nz.shape[0] > 0:
            rho[i] = nz.max()
        for _ in range(64):
            psum = 0.0
            for j in range(1, k):
                d = np.float32(dist[i, j] - rho[i])
                psum += np.exp(-(float(d) / mid)) if d > 0 else 1.0
            if abs(psum - target) < 1e-5:
                break
            if psum > target:
                hi = mid
                mid = (lo + hi) / 2.0
            else:
                lo = mid
                mid = mid * 2 if hi == np.inf else (lo + hi) / 2.0
        sig[i] = mid
        if rho[i] > 0.0:
            m = np.float32(dist[i].mean())
            if sig[i] < 1e-3 * m:
                sig[i] = 1e-3 * m
        elif sig[i] < 1e-3 * mean_all:
            sig[i] = 1e-3 * mean_all
    rows = np.repeat(np.arange(n), k)
    cols = knn_idx.reshape(-1)
    vals = np.zeros(n * k, np.float32)
    for i in range(n):
        for j in range(k):
            if knn_idx[i, j] == i:
                v = 0.0
            elif dist[i, j] - rho[i] <= 0.0 or sig[i] == 0.0:
                v = 1.0
            else:
                v = np.exp(-((dist[i, j] - rho[i]) / sig[i]))
            vals[i * k + j] = v
    res = sp.coo_matrix((vals, (rows, cols)), shape=(n, n))
    res.eliminate_zeros()
    tr = res.transpose()
    prod = res.multiply(tr)
    res = (res + tr - prod).tocsr()
    res.eliminate_zeros()
    res.sort_indices()
    return res


def synthetic_embedding(n: int, d: int = 128, n_clusters: int = 10, seed: int = 0) -> np.ndarray:
    """Z[N,d]: mixture of `n_clusters` unit-variance Gaussians, centres ~ N(0, 3²)."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, 3.0, size=(n_clusters, d))
    lab = rng.integers(0, n_clusters, size=n)
    return (centres[lab] + rng.normal(0.0, 1.0, size=(n, d))).astype(np.float32)


def synthetic_expression(n: int, g: int, density: float = 0.10, n_types: int = 10, seed: int = 0,
                         log_normalize: bool = True) -> np.ndarray:
    """Cell×gene matrix: NB counts (dispersion 0.5), log-normal gene means / size factors,
    `n_types` latent types shifting 5 % of the genes ×4, Bernoulli dropout to `density`."""
    rng = np.random.default_rng(seed)
    mu_g = rng.lognormal(0.0, 1.0, size=g)
    s_c = rng.lognormal(0.0, 0.5, size=n)
    types = rng.integers(0, n_types, size=n)
    shift = np.ones((n_types, g))
    for t in range(n_types):
        shift[t, rng.choice(g, size=max(1, g // 20), replace=False)] = 4.0
    out = np.empty((n, g), dtype=np.float32)
    r = 2.0  # NB "size" = 1/dispersion
    for i0 in range(0, n, 16384):
        i1 = min(n, i0 + 16384)
        mean = s_c[i0:i1, None] * mu_g[None, :] * shift[types[i0:i1]]
        lam = rng.gamma(shape=r, scale=mean / r)
        cnt = rng.poisson(lam).astype(np.float32)
        nz = (cnt > 0).mean()
        keep = min(1.0, density / max(nz, 1e-9))
        cnt *= rng.random(cnt.shape) < keep
        out[i0:i1] = cnt
    if log_normalize:
        out = log1p(normalize_total(out, target_sum=1e4))
    return out


# ----------------------------------------------------------------------------- f1: filters upstream of scGNN / GraphSCI
def scanpy_filter(x: np.ndarray, target: str, min_counts=None, min_other=None, max_counts=None, max_other=None):
    """``scanpy.pp.filter_genes`` / ``filter_cells`` (scanpy 1.10.1 `_simple.py`, un-vendored; published algorithm):
    number = X.sum(axis) for the *_counts criteria, (X > 0).sum(axis) for the *_cells / *_genes criteria; subset = number >= min
    or number <= max; exactly one criterion per call.  Called by the reference at dance/transforms/filter.py:121.  Unpinned
    (scanpy absent; the reference's own tests compare with scanpy itself, tests/transforms/test_filter_cell_gene.py)."""
    if sum(o is not None for o in (min_counts, min_other, max_counts, max_other)) != 1:
        raise ValueError("Only provide one of the optional parameters per call.")
    axis = 0 if target == "genes" else 1
    use_counts = min_counts is not None or max_counts is not None
    number = x.sum(axis) if use_counts else (x > 0).sum(axis)
    lo = min_counts if min_counts is not None else min_other
    hi = max_counts if max_counts is not None else max_other
    return (number >= lo) if lo is not None else (number <= hi), number


def gene_summary(x: np.ndarray, mode: str) -> np.ndarray:
    """FilterGenes summary statistic, dance/transforms/filter.py:480-489 (dense branch)."""
    if mode == "sum":
        return np.array(x.sum(0)).ravel()
    if mode == "var":
        return np.array((x**2).mean(0) - np.square(x.mean(0))).ravel()
    if mode == "cv":
        return np.nan_to_num(np.array(x.std(0) / x.mean(0)), posinf=0, neginf=0).ravel()
    if mode == "rv":
        return np.nan_to_num(np.array(x.var(0) / x.mean(0)), posinf=0, neginf=0).ravel()
    raise ValueError(mode)


def topk_gene_mask(summary: np.ndarray, num_genes: int, top: bool = True) -> np.ndarray:
    """FilterGenesTopK._get_preserve_mask, dance/transforms/filter.py:653-664."""
    num_genes = min(num_genes, summary.size)
    order = summary.argsort()
    mask = np.zeros(summary.size, dtype=bool)
    mask[order[-num_genes:] if top else order[:num_genes]] = True
    return mask
