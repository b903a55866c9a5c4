"""scDeepSort path (BASELINE config 0): CellFeatureGraph construction, AdaptiveSAGE aggregate, and the
ScDeepSort training loop against the oracle restatement (DGL-backed in the reference → restated)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _expr(n, g, seed):
    from oracle import port
    return port.synthetic_expression(n, g, density=0.15, seed=seed)


def test_cellgene_graph_matches_oracle(cuda):
    from dance_b200 import ops
    from oracle import port
    X = _expr(137, 61, 3)
    X[5] = 0            # an empty cell: no in-edges to renormalise
    X[:, 7] = 0         # an unexpressed gene
    src_r, dst_r, w_r = port.cell_feature_graph(X, normalize_edges=True)
    src, dst, w, nnz = ops.cellgene_graph(torch.from_numpy(X).to(cuda), True)
    assert nnz == int((X != 0).sum())
    assert torch.equal(src.cpu(), src_r) and torch.equal(dst.cpu(), dst_r)             # edge list and order: bit-exact
    assert np.allclose(w.cpu().numpy(), w_r.numpy(), rtol=2e-6, atol=0)                 # fp32 renormalised weights
    _, _, w_raw, _ = ops.cellgene_graph(torch.from_numpy(X).to(cuda), False)
    _, _, w_raw_r = port.cell_feature_graph(X, normalize_edges=False)
    assert torch.equal(w_raw.cpu(), w_raw_r)


def test_cellfeaturegraph_transform_and_subgraph(cuda):
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.transforms.graph import CellFeatureGraph
    X = _expr(50, 20, 1)
    rng = np.random.default_rng(0)
    ad = AnnDataLite(X, obsm={"pca": rng.normal(size=(50, 8)).astype(np.float32)}, varm={"pca": rng.normal(size=(20, 8)).astype(np.float32)})
    data = Data(ad, train_size=40)
    CellFeatureGraph(cell_feature_channel="pca")(data)
    g = data.data.uns["CellFeatureGraph"]
    assert g.number_of_nodes() == 70 and g.ndata["features"].shape == (70, 8)
    assert (g.ndata["cell_id"][:20] == torch.arange(20)).all() and (g.ndata["cell_id"][20:] == -1).all()
    nnz = int((X != 0).sum())
    assert g.num_edges() == 2 * nnz + 70
    # node-induced subgraph of genes ∪ train cells, as examples/…/scdeepsort.py:60-66 does
    sub = g.subgraph(torch.cat([torch.arange(20), torch.tensor(data.train_idx) + 20]))
    assert sub.number_of_nodes() == 60
    assert sub.num_edges() == 2 * int((X[:40] != 0).sum()) + 60
    import pickle
    g2 = pickle.loads(pickle.dumps(g))
    assert torch.equal(g2.src, g.src) and torch.equal(g2.edata["weight"], g.edata["weight"])


def test_adaptive_sage_neighbour_mean(cuda):
    from dance_b200 import ops
    from dance_b200.graph import GraphLite
    from dance_b200.modules.scdeepsort import ScDeepSort
    from oracle import port
    n, g, F = 90, 40, 16
    X = _expr(n, g, 5)
    src, dst, w = port.cell_feature_graph(X)
    h = torch.randn(n + g, F)
    alpha = torch.rand(g + 2, 1) + 0.5
    ref = port.adaptive_sage_neighbour_mean(src, dst, w, h, alpha, g)
    gr = GraphLite(src, dst, n + g)
    gr.edata["weight"] = w
    gr.ndata["features"] = h
    gr.ndata["cell_id"] = torch.cat([torch.arange(g, dtype=torch.int32), -torch.ones(n, dtype=torch.int32)])
    m = ScDeepSort(F, 8, 1, device="cuda")
    m._build(g, 3)
    m.params.p["alpha"].copy_(alpha)
    out = m.neighbour_mean(gr)
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_scdeepsort_training_matches_oracle(cuda, precision):
    """Three epochs of explicit mini-batches: loss curve, logits and weights ≤1e-4 rel vs the torch-CPU restatement."""
    from dance_b200.graph import GraphLite
    from dance_b200.modules.scdeepsort import ScDeepSort
    from oracle import port
    n, g, F, C = 600, 50, 400, 7
    rng = np.random.default_rng(1)
    feats = torch.from_numpy(rng.normal(size=(n + g, F)).astype(np.float32))
    labels = torch.from_numpy(rng.integers(0, C, size=n))
    X = _expr(n, g, 2)
    src, dst, w = port.cell_feature_graph(X)
    gr = GraphLite(src, dst, n + g)
    gr.edata["weight"] = w
    gr.ndata["features"] = feats
    gr.ndata["cell_id"] = torch.cat([torch.arange(g, dtype=torch.int32), -torch.ones(n, dtype=torch.int32)])
    cells = np.arange(g, g + n)
    batches = [[rng.permutation(cells)[i:i + 100] for i in range(0, n, 100)] for _ in range(3)]

    model = ScDeepSort(F, 200, 1, device="cuda", batch_size=100, precision=precision, seed=0)
    model._build(g, C)
    init = model.state_dict()
    net = port.ScDeepSortNet(F, 200, C, g)
    with torch.no_grad():
        net.sage_linear.weight.copy_(init["layers.0.layers.1.weight"].cpu()); net.sage_linear.bias.copy_(init["layers.0.layers.1.bias"].cpu())
        net.linear.weight.copy_(init["linear.weight"].cpu()); net.linear.bias.copy_(init["linear.bias"].cpu())
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
    full_labels = torch.cat([-torch.ones(g, dtype=torch.long), labels])
    ref_losses = [port.scdeepsort_epoch(net, opt, feats, full_labels, b) for b in batches]

    model._feat, model._lab = feats.to(cuda), full_labels.to(cuda)
    losses = [model.cal_loss(b, 1e-3, 1e-4) for b in batches]
    assert np.allclose(losses, ref_losses, rtol=1e-4)
    sd = model.state_dict()
    assert rel_err(sd["layers.0.layers.1.weight"].cpu().numpy(), net.sage_linear.weight.detach().numpy()) < 1e-4
    assert rel_err(sd["linear.weight"].cpu().numpy(), net.linear.weight.detach().numpy()) < 1e-4
    assert torch.equal(sd["alpha"].cpu(), torch.ones(g + 2, 1))                       # alpha never moves (no gradient in the reference)
    with torch.no_grad():
        ref_prob = torch.softmax(net(feats[g:]), -1).numpy()
    assert rel_err(model.predict_proba(gr), ref_prob) < 1e-4
    # full fit() API runs and improves on the training set
    model2 = ScDeepSort(F, 200, 1, device="cuda", batch_size=100, precision=precision, seed=0)
    model2.fit(gr, labels, epochs=3, lr=1e-3, weight_decay=0, val_ratio=0.2)
    assert model2.predict(gr).shape == (n, ) and model2.history[-1][0] < model2.history[0][0]


def test_scdeepsort_example_flow_end_to_end(cuda):
    """The flow of examples/single_modality/cell_type_annotation/scdeepsort.py:40-75 on synthetic data (BASELINE
    config 0 at reduced size): PCACellFeatureGraph + SetConfig pipeline → train/test subgraphs → fit → score."""
    from dance_b200.data import AnnDataLite, Data
    from dance_b200.modules.scdeepsort import ScDeepSort
    from dance_b200.transforms import Compose, PCACellFeatureGraph, SetConfig
    rng = np.random.default_rng(0)
    n, g, C = 1500, 300, 5
    types = rng.integers(0, C, size=n)
    base = rng.lognormal(0, 1, size=(C, g))
    X = rng.poisson(base[types] * rng.lognormal(0, 0.3, size=(n, 1))).astype(np.float32)
    X = np.log1p(X / X.sum(1, keepdims=True).clip(1) * 1e4).astype(np.float32)
    y = np.eye(C, dtype=np.float32)[types]
    data = Data(AnnDataLite(X, obsm={"cell_type": y}), train_size=1200)
    Compose(PCACellFeatureGraph(n_components=64, split_name="train"), SetConfig({"label_channel": "cell_type"}))(data)
    graph = data.data.uns["CellFeatureGraph"]
    y_all = data.get_y(return_type="torch")
    num_genes = data.shape[1]
    gene_ids = torch.arange(num_genes)
    train_ids = torch.tensor(data.train_idx) + num_genes
    test_ids = torch.tensor(data.test_idx) + num_genes
    g_train = graph.subgraph(torch.concat((gene_ids, train_ids)))
    g_test = graph.subgraph(torch.concat((gene_ids, test_ids)))
    model = ScDeepSort(64, 32, 1, "synthetic", "tissue", batch_size=200, device="cuda", seed=0)
    model.fit(g_train, y_all[data.train_idx].argmax(1), epochs=15, lr=1e-2, weight_decay=0, val_ratio=0.2)
    acc = model.score(g_test, y_all[data.test_idx].numpy())
    assert acc > 0.9, acc
    pred, unsure = model.predict(g_test, return_unsure=True)
    assert pred.shape == (300, ) and unsure.dtype == bool


@pytest.mark.parametrize("dropout", [0.0, 0.3])
def test_scdeepsort_two_layers_matches_torch_mlp(cuda, dropout):
    """GNN(n_layers = 2) (scdeepsort.py:26-88): every AdaptiveSAGE layer maps the destination node's own features (module
    docstring), so two layers are Linear-ReLU-Linear-ReLU-Linear on the cell features.  Two epochs of explicit mini-batches against
    torch autograd on the CPU with the same initial weights; with dropout the run is compared through its own recorded masks."""
    from dance_b200.modules.scdeepsort import ScDeepSort
    n, g, F, H, C = 400, 30, 64, 48, 5
    rng = np.random.default_rng(3)
    feats = torch.from_numpy(rng.normal(size=(n + g, F)).astype(np.float32))
    labels = torch.from_numpy(rng.integers(0, C, size=n))
    full_labels = torch.cat([-torch.ones(g, dtype=torch.long), labels])
    cells = np.arange(g, g + n)
    batches = [[rng.permutation(cells)[i:i + 100] for i in range(0, n, 100)] for _ in range(2)]
    model = ScDeepSort(F, H, 2, device="cuda", batch_size=100, precision="fp32", seed=0, dropout=dropout)
    model._build(g, C)
    sd = model.state_dict()
    assert "layers.1.layers.1.weight" in sd and sd["layers.1.layers.1.weight"].shape == (H, H) and "layers.1.alpha" in sd
    lin = [torch.nn.Linear(F, H), torch.nn.Linear(H, H), torch.nn.Linear(H, C)]
    with torch.no_grad():
        for i in (0, 1):
            lin[i].weight.copy_(sd[f"layers.{i}.layers.1.weight"].cpu()); lin[i].bias.copy_(sd[f"layers.{i}.layers.1.bias"].cpu())
        lin[2].weight.copy_(sd["linear.weight"].cpu()); lin[2].bias.copy_(sd["linear.bias"].cpu())
    opt = torch.optim.Adam([p for l in lin for p in l.parameters()], lr=1e-3)
    model._feat, model._lab = feats.to(cuda), full_labels.to(cuda)
    masks = []
    if dropout:                       # record the masks the model draws (same generator, same call order) for the reference
        gen = torch.Generator(device=cuda)
        gen.manual_seed(0)
        for ep in batches:
            for b in ep:
                masks.append([(torch.rand((len(b), w), device=cuda, generator=gen) >= dropout).float().cpu() / (1 - dropout) for w in (F, H)])
    losses, ref_losses, k = [], [], 0
    for ep in batches:
        losses.append(model.cal_loss(ep, 1e-3, 0.0))
        tot = cnt = 0.0
        for b in ep:
            x, y = feats[b], full_labels[b]
            m0, m1 = (masks[k] if dropout else (1.0, 1.0))
            k += 1
            opt.zero_grad()
            h = torch.relu(lin[1](torch.relu(lin[0](x * m0)) * m1))
            loss = torch.nn.functional.cross_entropy(lin[2](h), y, reduction="sum")
            loss.backward()
            opt.step()
            tot += loss.item() * len(b); cnt += len(b)
        ref_losses.append(tot / cnt)
    assert np.allclose(losses, ref_losses, rtol=1e-4)
    sd = model.state_dict()
    for i in (0, 1):
        assert rel_err(sd[f"layers.{i}.layers.1.weight"].cpu().numpy(), lin[i].weight.detach().numpy()) < 1e-4, i
    assert rel_err(sd["linear.weight"].cpu().numpy(), lin[2].weight.detach().numpy()) < 1e-4
