"""Diagnostics of the symmetric decoder at scale: NaN scan, sampled rows vs the fp64 closed form, agreement and timing vs the
row-sweep kernel.  python scripts/sym_check.py 200000 1000000"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from dance_b200 import ops  # noqa: E402
from test_gpu_kernels import gae_reference_rows  # noqa: E402

dev = torch.device("cuda:0")
for n in [int(a) for a in sys.argv[1:]] or [200_000]:
    gen = torch.Generator(device=dev).manual_seed(n)
    for scale in (0.1, 0.6):
        z = (torch.randn(n, 16, device=dev, generator=gen) * scale).contiguous()
        idx = torch.randint(0, n, (n, 8), device=dev, dtype=torch.int32, generator=gen)
        A = ops.knn_graph_build(idx.contiguous())
        L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
        rows = torch.cat([torch.tensor([0, 1, 127, 128, n - 1, n // 2], device=dev), torch.randint(0, n, (58, ), device=dev, generator=gen)])
        _, ref = gae_reference_rows(z, A.rowptr, A.colidx, 0.5, 100.0, rows)
        out = {}
        for path in ("sym", "f16"):
            ops.set_path("gae", path)
            loss, dz, _, _ = ops.gae_loss_grad(z, L, 0.5, 100.0)      # warm-up + result
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(2):
                ops.gae_loss_grad(z, L, 0.5, 100.0)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 2
            err = float((dz[rows].double() - ref).norm() / ref.norm())
            out[path] = (loss.item(), dz, ms)
            print(f"n={n} scale={scale} path={path}: {ms:8.2f} ms  loss={loss.item():.9g}  nan={int(torch.isnan(dz).sum())}  rows_vs_fp64={err:.3e}", flush=True)
        ops.set_path("gae", "auto")
        d = float((out['sym'][1] - out['f16'][1]).norm() / out['f16'][1].norm())
        print(f"   sym vs f16: dz rel {d:.3e}, loss rel {abs(out['sym'][0] - out['f16'][0]) / abs(out['f16'][0]):.3e}", flush=True)
