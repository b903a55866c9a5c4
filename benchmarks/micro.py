"""Kernel micro-benchmarks (CUDA-event timed on the launching stream, L2 flushed between
iterations).  Prints one JSON line per kernel with achieved algorithmic GB/s or TFLOP/s."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dance_b200 import ops  # noqa: E402

PEAKS = {"hbm_gbs": 6566.1, "bf16_tflops": 1746.9}
try:
    PEAKS.update(json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json")))
except Exception:
    pass


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def random_knn_graph(n, k, dev, seed=0, local=None):
    g = torch.Generator(device=dev).manual_seed(seed)
    if local is None:
        idx = torch.randint(0, n, (n, k), device=dev, dtype=torch.int32, generator=g)
    else:  # neighbours within a window of `local` rows (cluster-sorted cells)
        base = (torch.arange(n, device=dev) // local * local).unsqueeze(1)
        idx = (base + torch.randint(0, local, (n, k), device=dev, generator=g)).clamp_(max=n - 1).to(torch.int32)
    return ops.knn_graph_build(idx.contiguous())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    only = set(args.only.split(",")) if args.only else None
    out = []

    def want(name):
        return only is None or name in only

    n = args.n
    if want("spmm"):
        for local in (None, 100_000, 4096):
            A = random_knn_graph(n, 15, dev, local=local)
            for F in (16, 32, 64, 128):
                X = torch.randn(n, F, device=dev)
                Y = torch.empty(n, F, device=dev)
                med, best = timeit(lambda: ops.spmm(A, X, out=Y), flush=flush)
                alg = A.nnz * 8 + (n + 1) * 4 + 2 * n * F * 4
                gather = A.nnz * F * 4
                out.append(dict(kernel="spmm_csr_f32", n=n, nnz=A.nnz, F=F, locality=local, ms=med, ms_best=best,
                                alg_GB=alg / 1e9, alg_GBps=alg / med / 1e6, frac_hbm=alg / med / 1e6 / PEAKS["hbm_gbs"],
                                gather_GBps=gather / med / 1e6))
                print(json.dumps(out[-1]), flush=True)
    if want("gemm"):
        for prec in ("fp32", "tf32x3", "tf32"):
            for (M, N, K, tA, tB) in ((12800, 512, 2000, 0, 1), (12800, 2000, 512, 0, 1), (12800, 128, 512, 0, 1),
                                       (12800, 512, 2000, 0, 0), (2000, 512, 12800, 1, 0), (512, 2000, 12800, 1, 0)):
                A = torch.randn((K, M) if tA else (M, K), device=dev)
                B = torch.randn((N, K) if tB else (K, N), device=dev)
                C = torch.empty(M, N, device=dev)
                med, best = timeit(lambda: ops.gemm(A, B, transA=bool(tA), transB=bool(tB), out=C, precision=prec), flush=flush)
                fl = 2.0 * M * N * K
                out.append(dict(kernel="gemm_f32", precision=prec, M=M, N=N, K=K, tA=tA, tB=tB, ms=med, ms_best=best,
                                TFLOPs=fl / med / 1e9))
                print(json.dumps(out[-1]), flush=True)
    if want("gae"):
        for nn in (20_000, 100_000):
            A = random_knn_graph(nn, 15, dev)
            L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
            z = torch.randn(nn, 16, device=dev) * 0.3
            mu, lv = torch.randn(nn, 16, device=dev) * 0.1, torch.randn(nn, 16, device=dev) * 0.1
            med, best = timeit(lambda: ops.gae_loss_grad(z, L, 0.5, 100.0, mu, lv), iters=3, warmup=1)
            out.append(dict(kernel="gae_loss_grad", n=nn, ms=med, pairs_per_s=nn * nn / med * 1e3))
            print(json.dumps(out[-1]), flush=True)
    if want("knn"):
        for nn, d in ((100_000, 128), (100_000, 50)):
            X = torch.randn(nn, d, device=dev) + torch.randn(10, d, device=dev)[torch.randint(0, 10, (nn, ), device=dev)] * 3
            med, best = timeit(lambda: ops.knn(X, 15, return_dist=False), iters=3, warmup=1)
            out.append(dict(kernel="knn_l2", n=nn, d=d, k=15, ms=med, TFLOPs=2.0 * nn * nn * d / med / 1e9))
            print(json.dumps(out[-1]), flush=True)
    if want("normalize"):
        for nn, g in ((200_000, 2000), ):
            X = torch.rand(nn, g, device=dev)
            med, best = timeit(lambda: ops.normalize_total_log1p_(X, target_sum=1e4), flush=flush)
            alg = 2.0 * nn * g * 4
            out.append(dict(kernel="normalize_total_log1p", n=nn, g=g, ms=med, alg_GBps=alg / med / 1e6,
                            frac_hbm=alg / med / 1e6 / PEAKS["hbm_gbs"]))
            print(json.dumps(out[-1]), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    json.dump(out, open("gpurun_out/micro.json", "w"), indent=1)


if __name__ == "__main__":
    main()
