"""One decoder call (profiling target): python scripts/sym_one.py <n_cells> [path]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dance_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
path = sys.argv[2] if len(sys.argv) > 2 else "auto"
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
z = (torch.randn(n, 16, device=dev, generator=gen) * 0.3).contiguous()
idx = torch.randint(0, n, (n, 8), device=dev, dtype=torch.int32, generator=gen)
A = ops.knn_graph_build(idx.contiguous())
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
ops.set_path("gae", path)
for _ in range(2):
    loss, dz, _, _ = ops.gae_loss_grad(z, L, 0.5, 100.0)
torch.cuda.synchronize()
print(n, path, loss.item())
