"""Accuracy of the GEMM paths vs fp64 for growing K (diagnostic)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dance_b200 import ops

rng = np.random.default_rng(0)
for (M, N, K) in ((256, 128, 64), (256, 128, 256), (256, 128, 1024), (256, 128, 4096), (1000, 512, 2000), (1000, 2000, 512), (12800, 512, 2000)):
    A = rng.normal(size=(M, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    out = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        C = ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), transB=True, precision=prec).cpu().numpy()
        out[prec] = (np.linalg.norm(C - ref) / np.linalg.norm(ref), np.abs(C - ref).max() / np.abs(ref).max())
    tc = (torch.from_numpy(A).cuda() @ torch.from_numpy(B).cuda().t()).cpu().numpy()
    out["torch"] = (np.linalg.norm(tc - ref) / np.linalg.norm(ref), 0)
    print(M, N, K, {k: "%.2e/%.2e" % v for k, v in out.items()}, flush=True)
# positive-only inputs (ReLU activations): truncation bias shows up here
for K in (512, 2000):
    A = np.abs(rng.normal(size=(512, K))).astype(np.float32)
    B = np.abs(rng.normal(size=(256, K))).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    for prec in ("fp32", "tf32x3", "tf32"):
        C = ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), transB=True, precision=prec).cpu().numpy()
        print("positive", K, prec, "rel %.2e mean signed rel %.2e" % (np.linalg.norm(C - ref) / np.linalg.norm(ref), ((C - ref) / ref).mean()), flush=True)
