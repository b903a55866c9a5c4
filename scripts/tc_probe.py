"""Bring-up probe for the MN-major tcgen05 descriptor fields: tries the candidate encodings in
separate processes (env overrides read by gemm_tc()) and prints the relative error of each."""
import itertools
import os
import subprocess
import sys

CHILD = r'''
import numpy as np, torch, sys
sys.path.insert(0, ".")
from dance_b200 import ops
rng = np.random.default_rng(0)
res = []
for (tA, tB, M, N, K) in ((0,0,256,128,64), (1,1,256,128,64), (1,0,256,256,96)):
    A = rng.normal(size=(K, M) if tA else (M, K)).astype(np.float32)
    B = rng.normal(size=(N, K) if tB else (K, N)).astype(np.float32)
    ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    C = ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), transA=bool(tA), transB=bool(tB), precision="tf32")
    torch.cuda.synchronize()
    res.append(float(np.linalg.norm(C.cpu().numpy() - ref) / np.linalg.norm(ref)))
print("ERR", res)
'''

combos = list(itertools.product((1, 2), (4, 3), (512, 1024), (4096, )))  # layout, swizzle enum, sbo, lbo
combos += [(1, 4, 4096, 512), (2, 3, 4096, 1024)]                        # swapped LBO/SBO
for layout, swz, sbo, lbo in combos:
    env = dict(os.environ, B2_TC_MN_LAYOUT=str(layout), B2_TC_MN_SWZ=str(swz), B2_TC_MN_SBO=str(sbo), B2_TC_MN_LBO=str(lbo))
    try:
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=60)
        line = [l for l in out.stdout.splitlines() if l.startswith("ERR")]
        print(f"layout={layout} swz={swz} sbo={sbo} lbo={lbo} ->", line[0] if line else ("FAIL " + out.stderr[-300:]), flush=True)
    except subprocess.TimeoutExpired:
        print(f"layout={layout} swz={swz} sbo={sbo} lbo={lbo} -> TIMEOUT", flush=True)
