"""Timeline of CTA 0 of the symmetric decoder (lab build with -DB2_GAE_TRACE): python scripts/lab/sym_trace.py [n] [tiles]
Writes gpurun_out/sym_trace_<mode>.npy ([5 agents, tiles, 8 stamps], cycles relative to the first stamp) and prints a summary."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from dance_b200 import _lib  # noqa: E402

_lib._LIB_PATH = ROOT / "scripts" / "lab" / "libdance_b200_trace.so"
from dance_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
z = (torch.randn(n, 16, device=dev, generator=gen) * 0.3).contiguous()
idx = torch.randint(0, n, (n, 8), device=dev, dtype=torch.int32, generator=gen)
A = ops.knn_graph_build(idx.contiguous())
L = ops.CSR(A.rowptr, A.colidx, None, A.shape)
h = _lib.lib()
h.b2_debug_gae_sym_trace.restype = C.c_int
h.b2_debug_gae_sym_trace.argtypes = [C.c_void_p, C.c_int]
buf = torch.zeros(5 * T * 8, dtype=torch.int64, device=dev)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
for mode, knobs in (("early_gwait", {"gae_late_gempty": 0}), ("late_gwait", {"gae_late_gempty": 1})):
    for k, v in knobs.items():
        ops.set_tuning(k, v)
    ops.gae_loss_grad(z, L, 0.5, 100.0)
    torch.cuda.synchronize()
    buf.zero_()
    h.b2_debug_gae_sym_trace(buf.data_ptr(), T)
    ops.gae_loss_grad(z, L, 0.5, 100.0)
    torch.cuda.synchronize()
    h.b2_debug_gae_sym_trace(None, 0)
    tr = buf.cpu().numpy().reshape(5, T, 8).astype(np.int64)
    t0 = tr[tr > 0].min()
    rel = np.where(tr > 0, tr - t0, -1)
    np.save(ROOT / "gpurun_out" / f"sym_trace_{mode}.npy", rel)
    # steady-state summary over tiles 32..T/2-1 of each group (each group sees every other tile)
    lo, hi = 16, T // 2 - 8
    print(f"== {mode}: cycles per tile of one group (mean over its tiles {lo}..{hi})")
    for q in (0, 1):
        e = rel[q, lo:hi]
        cyc = np.diff(e[:, 5]).mean()
        print(f"  EW group {q}: period {cyc:8.0f} | wait S {np.mean(e[:, 1] - e[:, 0]):7.0f} | wait G-buffer {np.mean(e[:, 2] - e[:, 1]):7.0f} | "
              f"first half {np.mean(e[:, 3] - e[:, 2]):7.0f} | second half {np.mean(e[:, 4] - e[:, 3]):7.0f} | fence+arrive {np.mean(e[:, 5] - e[:, 4]):6.0f}")
    s_ = rel[2, 2 * lo:2 * hi]
    d_ = rel[3, 2 * lo:2 * hi]
    print(f"  issuer S: wait s_empty {np.mean(s_[:, 1] - s_[:, 0]):7.0f} | wait Z stage {np.mean(s_[:, 2] - s_[:, 1]):6.0f} | issue {np.mean(s_[:, 3] - s_[:, 2]):6.0f} | "
          f"bookkeeping {np.mean(s_[1:, 0] - s_[:-1, 3]):6.0f}")
    print(f"  issuer D: wait g_full {np.mean(d_[:, 1] - d_[:, 0]):7.0f} | accumulator waits + issue {np.mean(d_[:, 2] - d_[:, 1]):6.0f} | commits {np.mean(d_[:, 3] - d_[:, 2]):6.0f} | "
          f"bookkeeping {np.mean(d_[1:, 0] - d_[:-1, 3]):6.0f} | tile period {np.diff(d_[:, 2]).mean():7.0f}")
    # lag between EW(k) end and D(k) issue complete; and between D(k) issued and the group's next g_empty pass
    k = np.arange(2 * lo, 2 * hi)
    ew_end = np.array([rel[kk & 1, kk >> 1, 5] for kk in k])
    ew_next_gempty = np.array([rel[kk & 1, (kk >> 1) + 1, 2] for kk in k])
    print(f"  EW(k) end → D(k) issued {np.mean(d_[:, 2] - ew_end):6.0f} | D(k) issued → group passes its G-buffer wait {np.mean(ew_next_gempty - d_[:, 2]):6.0f}")
    fl = rel[4, lo:hi]
    print(f"  flush: wait d2_full {np.mean(fl[:, 1] - fl[:, 0]):7.0f} | drain+reds {np.mean(fl[:, 2] - fl[:, 1]):6.0f}")
    # first 12 tiles, raw
    for kk in range(24, 36):
        q, i = kk & 1, kk >> 1
        print(f"   tile {kk:3d} g{q}: EW {rel[q, i, :6].tolist()}  S {rel[2, kk, :4].tolist()}  D {rel[3, kk, :4].tolist()}")
