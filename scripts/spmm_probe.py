"""Aggregate (CSR SpMM) probe on a scGNN-shaped graph: python scripts/spmm_probe.py [n_cells] [F] [reps] [variants]
Builds the symmetrised kNN graph (k = 15) of a clustered 128-d embedding (10 types), then times each variant with CUDA events,
flushing L2 (256 MB write) between launches.  variants: comma list of f32,bf16,f16 (+ experimental names the library exports)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dance_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
variants = (sys.argv[4] if len(sys.argv) > 4 else "f32,bf16").split(",")
order = sys.argv[5] if len(sys.argv) > 5 else "locality"
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
types = torch.randint(0, 10, (n, ), device=dev, generator=gen)
centers = torch.randn(10, 128, device=dev, generator=gen)
emb = (centers[types] + 0.7 * torch.randn(n, 128, device=dev, generator=gen)).relu_().contiguous()
if order == "locality":
    perm, inv = ops.locality_order(emb, n_anchors=64)
    emb = emb[perm].contiguous()
idx, _ = ops.knn(emb, 15, return_dist=False)
A = ops.knn_graph_build(idx.contiguous())
del emb
nnz = A.nnz
if len(sys.argv) > 6:        # dump the CSR for scripts/lab/gather_lab (int64 n, int64 nnz, rowptr, colidx, vals)
    import numpy as np
    with open(sys.argv[6], "wb") as f:
        np.array([n, nnz], dtype=np.int64).tofile(f)
        A.rowptr.cpu().numpy().astype(np.int32).tofile(f)
        A.colidx[:nnz].cpu().numpy().astype(np.int32).tofile(f)
        A.vals[:nnz].cpu().numpy().astype(np.float32).tofile(f)
    print("dumped", sys.argv[6], flush=True)
X = torch.randn(n, F, device=dev, generator=gen).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for v in variants:
    if v == "f32":
        Xv, kw = X, {}
    elif v in ("bf16", "f16"):
        Xv, kw = ops.to_x16(X, torch.bfloat16 if v == "bf16" else torch.float16), {}
    elif v.startswith("stage"):          # stage32 / stage16: staged-gather kernels (ops.spmm_staged)
        Xv = X if v == "stage32" else ops.to_x16(X, torch.bfloat16)
        kw = {"staged": True}
    else:
        raise SystemExit(f"unknown variant {v}")
    out = torch.empty(n, F, device=dev)
    fn = (lambda: ops.spmm_staged(A, Xv, out=out)) if kw.get("staged") else (lambda: ops.spmm(A, Xv, out=out))
    fn()
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
        err = 0.0
    else:
        err = ((out - ref).norm() / ref.norm()).item()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        fn()
    e.record()
    torch.cuda.synchronize()
    warm = s.elapsed_time(e) / 10
    esz = 4 if Xv.dtype == torch.float32 else 2
    alg = nnz * 8 + (n + 1) * 4 + n * F * esz + n * F * 4
    t = sorted(ts)[len(ts) // 2]
    print(f"n={n} F={F} nnz={nnz} {v:8s} cold {t:.3f} ms ({alg / t / 1e6:.0f} GB/s, {alg / t / 1e6 / 6566.1 * 100:.1f} % of HBM peak)  "
          f"back-to-back {warm:.3f} ms   rel diff vs first variant {err:.2e}", flush=True)
