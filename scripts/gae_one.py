"""One decoder call at n cells (for ncu): python scripts/gae_one.py [n]"""
import sys, torch
sys.path.insert(0, ".")
from dance_b200 import ops
import benchmarks.micro as m
dev = torch.device("cuda:0")
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
A = m.random_knn_graph(nn, 15, dev)
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
z = torch.randn(nn, 16, device=dev) * 0.3
for _ in range(2):
    loss, dz, _, _ = ops.gae_loss_grad(z, L, 0.5, 100.0)
torch.cuda.synchronize()
print(loss.item())
