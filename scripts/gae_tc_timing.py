"""Timing experiments for the tcgen05 decoder (wrong results in debug modes, only the time matters)."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from dance_b200 import ops
import benchmarks.micro as m
dev = torch.device("cuda:0")
nn = 100000
A = m.random_knn_graph(nn, 15, dev)
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
z = torch.randn(nn, 16, device=dev) * 0.3
med, best = m.timeit(lambda: ops.gae_loss_grad(z, L, 0.5, 100.0), iters=5, warmup=2)
print("MS", med)
'''
for mode in (0, 2, 4, 7):
    env = dict(os.environ, B2_GAE_TC_DEBUG=str(mode))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print("debug", mode, [l for l in out.stdout.splitlines() if l.startswith("MS")] or out.stderr[-300:], flush=True)
