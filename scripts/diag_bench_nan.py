"""Where does the Graph-AE of bench.py go non-finite?  Prints finiteness / magnitudes after every stage, for a few variants."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from dance_b200 import ops  # noqa: E402
from dance_b200.engine import FeatureAEEngine, GraphAEEngine  # noqa: E402

dev = torch.device("cuda:0")


def stat(name, t):
    t = t.float()
    print(f"    {name:14s} finite={bool(torch.isfinite(t).all())} absmax={float(t.abs().max()):.4g} mean={float(t.mean()):.4g}", flush=True)


def run(N, order, path, steps=3):
    print(f"== N={N} order={order} path={path}", flush=True)
    ops.set_path("gae", path)
    X = bench.synth_expression(N, 2000, dev)
    stat("X", X)
    fae = FeatureAEEngine(2000, device=dev, lr=1e-3, precision="tf32x3", seed=0)
    gae = GraphAEEngine(128, 16, device=dev, lr=1e-2, precision="tf32x3", seed=1)
    z_all = torch.empty(N, 128, device=dev)
    loss = fae.train_epoch(X, 12800, "LTMG", 0.9, None, z_all, None)
    print(f"    fae loss/cell = {loss.item() / N:.5g}", flush=True)
    stat("z_all", z_all)
    idx, _ = ops.knn(z_all, 15, include_rank0=False, return_dist=False)
    perm = None
    if order == "locality":
        perm, inv = ops.locality_order(z_all, n_anchors=64)
        idx = inv[idx[perm].long()].to(torch.int32)
        print(f"    perm ok={bool(torch.equal(inv[perm], torch.arange(N, device=dev)))} idx range=({int(idx.min())},{int(idx.max())})", flush=True)
    A = ops.knn_graph_build(idx.contiguous())
    labels = ops.CSR(A.rowptr, A.colidx, None, A.shape)
    adj_sum = A.nnz - N
    pw, norm = float(N * N - adj_sum) / adj_sum, N * N / float((N * N - adj_sum) * 2)
    print(f"    nnz={A.nnz} pw={pw:.4g} norm={norm:.6g}", flush=True)
    stat("A.vals", A.vals)
    gen = torch.Generator(device=dev).manual_seed(99)
    eps = torch.empty(N, 16, device=dev)
    for it in range(steps):
        fae.train_epoch(X, 12800, "LTMG", 0.9, None, z_all, None)
        zin = z_all if perm is None else z_all[perm].contiguous()
        eps.normal_(generator=gen)
        z, mu, lv = gae.train_step(zin, A, labels, norm, pw, eps)
        b = gae._buffers(N)
        print(f"  step {it}: loss={gae.loss.item():.6g}", flush=True)
        for nm, t in (("zin", zin), ("mu", mu), ("logvar", lv), ("z", z), ("dz", b["dz"]), ("dml", b["dml"]), ("grad", gae.params.grad), ("w", gae.params.flat)):
            stat(nm, t)
        if not torch.isfinite(gae.params.flat).all():
            break
    ops.set_path("gae", "auto")
    del X, fae, gae
    torch.cuda.empty_cache()


variants = [(200_000, "data", "auto"), (200_000, "locality", "auto"), (1_000_000, "data", "f16"), (1_000_000, "data", "auto")]
for v in variants:
    try:
        run(*v)
    except Exception as ex:
        print("   EXC", type(ex).__name__, ex, flush=True)
