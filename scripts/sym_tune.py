"""Timing of the symmetric decoder under its scheduling knobs.  python scripts/sym_tune.py 200000"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dance_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
z = (torch.randn(n, 16, device=dev, generator=gen) * 0.3).contiguous()
idx = torch.randint(0, n, (n, 8), device=dev, dtype=torch.int32, generator=gen)
A = ops.knn_graph_build(idx.contiguous())
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)


def t():
    ops.gae_loss_grad(z, L, 0.5, 100.0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        ops.gae_loss_grad(z, L, 0.5, 100.0)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / 3


for late in (0, 1):
    ops.set_tuning("gae_late_gempty", late)
    for stag in (0, 1500):
        ops.set_tuning("gae_stagger", stag)
        print(f"n={n} late_gempty={late} stagger={stag}: {t():.2f} ms", flush=True)
ops.set_tuning("gae_late_gempty", 0)
ops.set_tuning("gae_stagger", 1500)
# large-magnitude embedding: the scaled-operand variant
zb = z * 3.0e4
loss, dz, _, _ = ops.gae_loss_grad(zb, L, 0.5, 100.0)
print("large |z|: finite", bool(torch.isfinite(dz).all()), float(loss.item()), f"{t():.2f} ms (unscaled input timing)")
