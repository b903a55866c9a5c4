"""Diagnose the tf32x3 error growth with problem size: determinism + error vs M + distance to the emulated 3-term sum."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dance_b200 import ops

torch.manual_seed(0)
N, K = 512, 2000
B = torch.randn(N, K, device="cuda")
Bh = (B.view(torch.int32) & -8192).view(torch.float32)
Bl = B - Bh
for M in (128, 256, 1024, 4096, 12800, 51200):
    A = torch.randn(M, K, device="cuda")
    Ah = (A.view(torch.int32) & -8192).view(torch.float32)
    Al = A - Ah
    ref = A.double() @ B.double().t()
    emu = Ah.double() @ Bh.double().t() + Ah.double() @ Bl.double().t() + Al.double() @ Bh.double().t()
    C1 = ops.gemm(A, B, transB=True, precision="tf32x3")
    C2 = ops.gemm(A, B, transB=True, precision="tf32x3")
    torch.cuda.synchronize()
    rel = lambda x, y: ((x.double() - y).norm() / y.norm()).item()
    # per row-tile error to see whether late tiles of a persistent CTA are worse
    tile_err = [(rel(C1[i:i + 128], ref[i:i + 128])) for i in range(0, M, 128)]
    print(f"M={M}: rel_vs_fp64={rel(C1, ref):.2e} rel_vs_emulated={rel(C1, emu):.2e} emu_vs_fp64={rel(emu.float(), ref):.2e} "
          f"bitwise_repeatable={bool(torch.equal(C1, C2))} tile_err[min/med/max]={min(tile_err):.1e}/{np.median(tile_err):.1e}/{max(tile_err):.1e}",
          flush=True)
