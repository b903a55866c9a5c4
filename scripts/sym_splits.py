"""Decoder time at n cells for forced step splits: python scripts/sym_splits.py 1000000 1,2,3"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dance_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
splits = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2").split(",")]
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
z = (torch.randn(n, 16, device=dev, generator=gen) * 0.3).contiguous()
idx = torch.randint(0, n, (n, 8), device=dev, dtype=torch.int32, generator=gen)
A = ops.knn_graph_build(idx.contiguous())
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
for sp in splits:
    ops.set_tuning("gae_splits", sp)
    ops.gae_loss_grad(z, L, 0.5, 100.0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(2):
        loss, dz, _, _ = ops.gae_loss_grad(z, L, 0.5, 100.0)
    e.record()
    torch.cuda.synchronize()
    print(f"n={n} splits={sp} (0 = automatic): {s.elapsed_time(e) / 2:.2f} ms  loss {loss.item():.9f}", flush=True)
