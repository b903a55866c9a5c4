import sys, numpy as np, torch
sys.path.insert(0, ".")
from dance_b200 import ops
for g in (64, 128, 255, 256, 257, 301, 512):
    rng = np.random.default_rng(g)
    A = rng.normal(size=(g + 20, g)).astype(np.float32)
    Cm = (A.T @ A).astype(np.float32)
    ev, V, sweeps = ops.sym_eig(torch.from_numpy(Cm.copy()).cuda())
    ev, V = ev.cpu().numpy().astype(np.float64), V.cpu().numpy().astype(np.float64)
    ref = np.linalg.eigvalsh(Cm.astype(np.float64))[::-1]
    print(g, "sweeps", sweeps, "max|ev-ref|/ref0 %.2e" % (np.abs(ev - ref).max() / ref[0]), "orth %.2e" % np.abs(V @ V.T - np.eye(g)).max(),
          "diag %.2e" % (np.linalg.norm(V @ Cm @ V.T - np.diag(ev)) / np.linalg.norm(ev)), flush=True)
