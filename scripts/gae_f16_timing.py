"""Decoder timing: fp16-split tcgen05 (default) vs tf32-split (B2_GAE_NO_F16=1) vs CUDA cores (B2_GAE_NO_TC=1)."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from dance_b200 import ops
import benchmarks.micro as m
dev = torch.device("cuda:0")
nn = int(sys.argv[1])
A = m.random_knn_graph(nn, 15, dev)
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
z = torch.randn(nn, 16, device=dev) * 0.3
med, best = m.timeit(lambda: ops.gae_loss_grad(z, L, 0.5, 100.0), iters=5, warmup=2)
loss, dz, _, _ = ops.gae_loss_grad(z, L, 0.5, 100.0)
print("MS", med, "loss", loss.item(), "dz", dz.double().norm().item())
'''
for nn in (100000, 200000):
    for name, env in (("f16", {}), ("f16 packed-B dZ (8 MMAs)", {"B2_GAE_PACKED": "1"}), ("tf32", {"B2_GAE_NO_F16": "1"}), ("cuda-core", {"B2_GAE_NO_TC": "1"}),
                      ("f16 no-SFU", {"B2_GAE_TC_DEBUG": "1"}), ("f16 no-dZ", {"B2_GAE_TC_DEBUG": "2"}), ("f16 no-S", {"B2_GAE_TC_DEBUG": "4"}),
                      ("f16 none", {"B2_GAE_TC_DEBUG": "7"})):
        if nn > 100000 and "no-" in name or (nn > 100000 and name == "f16 none"):
            continue
        out = subprocess.run([sys.executable, "-c", CHILD, str(nn)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        print(nn, name, [l for l in out.stdout.splitlines() if l.startswith("MS")] or out.stderr[-300:], flush=True)
