"""GraphConvLayer (a13): dgl.nn.GraphConv(norm="both") and graph-sc's WeightedGraphConv, forward + backward against torch
restatements (oracle/dgl_lite.py for the former; the in-tree forward of graphsc.py:428-484 restated below)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _graph(n, deg, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, size=n * deg)
    dst = rng.integers(0, n, size=n * deg)
    src = np.concatenate([src, np.arange(n)])          # self loops: no zero in-degree
    dst = np.concatenate([dst, np.arange(n)])
    key = np.unique(src * n + dst)
    return (key // n).astype(np.int64), (key % n).astype(np.int64)


@pytest.mark.parametrize("fin,fout,act", [(300, 64, "tanh"), (48, 48, "relu"), (32, 120, None)])
def test_dgl_graphconv_forward_backward(cuda, fin, fout, act):
    from dance_b200.layers import GraphConvLayer
    from oracle import dgl_lite
    n = 500
    src, dst = _graph(n, 6, fin)
    g = dgl_lite.Graph(src, dst, n)
    ref = dgl_lite.GraphConv(fin, fout, activation={"tanh": torch.tanh, "relu": torch.relu, None: None}[act])
    with torch.no_grad():
        ref.bias.copy_(torch.randn(fout) * 0.1)
    x = torch.randn(n, fin, requires_grad=True)
    y = ref(g, x)
    up = torch.randn(n, fout)
    y.backward(up)
    lay = GraphConvLayer(fin, fout, activation=act, device=cuda)
    lay.weight.copy_(ref.weight.detach())
    lay.bias.copy_(ref.bias.detach())
    lay.bind(src, dst, n)
    assert lay.weight_first == (fin > fout)
    out = lay(x.detach().to(cuda))
    assert rel_err(out, y.detach()) < 1e-5
    dx = lay.backward(up.to(cuda))
    assert rel_err(dx, x.grad) < 1e-4 and rel_err(lay.grad_weight, ref.weight.grad) < 1e-4 and rel_err(lay.grad_bias, ref.bias.grad) < 1e-4
    with pytest.raises(RuntimeError):
        GraphConvLayer(4, 4, device=cuda).bind([0], [1], 3)        # zero in-degree nodes, as dgl


@pytest.mark.parametrize("norm,agg", [("both", "sum"), ("right", "sum"), ("none", "mean")])
def test_weighted_graphconv_matches_in_tree_forward(cuda, norm, agg):
    from dance_b200.layers import GraphConvLayer
    n, fin, fout = 400, 50, 20
    src, dst = _graph(n, 5, 3)
    rng = np.random.default_rng(0)
    w_e = rng.uniform(0.1, 2.0, size=len(src)).astype(np.float32)
    W = torch.randn(fin, fout, requires_grad=True)
    b = torch.randn(fout, requires_grad=True)
    x = torch.randn(n, fin, requires_grad=True)
    s, d = torch.from_numpy(src), torch.from_numpy(dst)
    indeg = torch.bincount(d, minlength=n).float().clamp(min=1)
    outdeg = torch.bincount(s, minlength=n).float().clamp(min=1)
    h = x
    if norm == "both":                                           # graphsc.py:446-451
        h = h * outdeg.pow(-0.5)[:, None]
    h = h @ W                                                    # :461-462 (weight always first)
    m = h[s] * torch.from_numpy(w_e)[:, None]                    # edge_selection_simple :418-426
    rst = torch.zeros(n, fout).index_add(0, d, m)
    if agg == "mean":                                            # fn.mean :466-467
        rst = rst / indeg[:, None]
    if norm != "none":                                           # :469-477
        rst = rst * (indeg.pow(-0.5) if norm == "both" else 1.0 / indeg)[:, None]
    y = torch.relu(rst + b)
    up = torch.randn(n, fout)
    y.backward(up)
    lay = GraphConvLayer(fin, fout, norm=norm, agg=agg, edge_weighted=True, weight_first=True, activation="relu", device=cuda)
    lay.weight.copy_(W.detach())
    lay.bias.copy_(b.detach())
    lay.bind(src, dst, n, edge_weight=w_e)
    assert rel_err(lay(x.detach().to(cuda)), y.detach()) < 1e-5
    dx = lay.backward(up.to(cuda))
    assert rel_err(dx, x.grad) < 1e-4 and rel_err(lay.grad_weight, W.grad) < 1e-4 and rel_err(lay.grad_bias, b.grad) < 1e-4


@pytest.mark.parametrize("k,act,weighted", [(2, None, True), (3, "relu", False), (1, None, True)])
def test_tagconv_forward_backward(cuda, k, act, weighted):
    """scTAG's dgl TAGConv (sctag.py:101-102,173-174): hop stack + one Linear, forward and every gradient vs the dgl_lite restatement."""
    from dance_b200.layers import TAGConvLayer
    from oracle import dgl_lite
    n, fin, fout = 600, 48, 20
    src, dst = _graph(n, 5, k)
    rng = np.random.default_rng(1)
    w_e = rng.uniform(0.2, 1.5, size=len(src)).astype(np.float32) if weighted else None
    g = dgl_lite.Graph(src, dst, n)
    ref = dgl_lite.TAGConv(fin, fout, k=k, activation={"relu": torch.relu, None: None}[act])
    with torch.no_grad():
        ref.lin.bias.copy_(torch.randn(fout) * 0.1)
    x = torch.randn(n, fin, requires_grad=True)
    y = ref(g, x, edge_weight=None if w_e is None else torch.from_numpy(w_e))
    up = torch.randn(n, fout)
    y.backward(up)
    lay = TAGConvLayer(fin, fout, k=k, activation=act, device=cuda)
    lay.weight.copy_(ref.lin.weight.detach())
    lay.bias.copy_(ref.lin.bias.detach())
    lay.bind(src, dst, n, edge_weight=w_e)
    assert rel_err(lay(x.detach().to(cuda)), y.detach()) < 1e-5
    dx = lay.backward(up.to(cuda))
    assert rel_err(dx, x.grad) < 1e-4 and rel_err(lay.grad_weight, ref.lin.weight.grad) < 1e-4 and rel_err(lay.grad_bias, ref.lin.bias.grad) < 1e-4


@pytest.mark.parametrize("flavour", ["scdsc", "dstg", "stdgcn"])
def test_in_tree_gcn_layers(cuda, flavour):
    """scDSC GNNLayer (scdsc.py:494-500), DSTG GraphConvolution (dstg.py:75-97), STdGCN conGraphConvolutionlayer (stdgcn.py:82-88):
    ``spmm(adj, x @ W) (+ bias) (relu)`` — forward and gradients against the same three torch lines the reference runs."""
    import scipy.sparse as sp
    from dance_b200.layers import AdjLinearLayer
    n, fin, fout = 700, 64, 24
    src, dst = _graph(n, 6, 9)
    rng = np.random.default_rng(2)
    A = sp.csr_matrix((rng.uniform(0.05, 1.0, size=len(src)).astype(np.float32), (dst, src)), shape=(n, n))
    bias = flavour != "scdsc"
    act = "relu" if flavour == "scdsc" else None
    lay = AdjLinearLayer(fin, fout, bias=bias, activation=act, init="uniform_out" if flavour == "stdgcn" else "xavier", device=cuda, seed=0)
    lay.bind(A)
    W = lay.weight.detach().cpu().clone().requires_grad_()
    b = lay.bias.detach().cpu().clone().requires_grad_() if bias else None
    x = torch.randn(n, fin, requires_grad=True)
    adj_t = torch.sparse_coo_tensor(np.vstack(A.nonzero()), A.data, (n, n)).coalesce()
    out = torch.spmm(adj_t, torch.mm(x, W))
    if bias:
        out = out + b
    if act:
        out = torch.relu(out)
    up = torch.randn(n, fout)
    out.backward(up)
    y = lay(x.detach().to(cuda))
    assert rel_err(y, out.detach()) < 1e-5
    dx = lay.backward(up.to(cuda))
    assert rel_err(dx, x.grad) < 1e-4 and rel_err(lay.grad_weight, W.grad) < 1e-4
    if bias:
        assert rel_err(lay.grad_bias, b.grad) < 1e-4
    if flavour == "scdsc":          # `active=False` switches the ReLU off (scdsc.py:497-499)
        assert rel_err(lay(x.detach().to(cuda), active=False), torch.spmm(adj_t, torch.mm(x, W)).detach()) < 1e-5
