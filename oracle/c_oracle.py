"""ctypes loader of the plain-C restatements in ``oracle/c`` (built by ``oracle/Makefile`` into ``oracle/_build``).
TEST INFRASTRUCTURE: an independent checker for ``oracle/port.py``; never imported by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = _HERE / "_build" / "liboracle_c.so"
        srcs = list((_HERE / "c").glob("*.c"))
        if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
            subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)
        _lib = C.CDLL(str(so))
        _lib.knn_rank_f64.restype = C.c_int
        _lib.gae_loss_grad_f64.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def csr_spmm(indptr, indices, vals, X):
    X = np.ascontiguousarray(X, np.float32)
    Y = np.empty((len(indptr) - 1, X.shape[1]), np.float32)
    indptr, indices = np.ascontiguousarray(indptr, np.int32), np.ascontiguousarray(indices, np.int32)
    vals = None if vals is None else np.ascontiguousarray(vals, np.float32)
    lib().csr_spmm_f32(_p(indptr), _p(indices), _p(vals) if vals is not None else None, _p(X), C.c_int64(X.shape[1]), _p(Y),
                       C.c_int64(X.shape[1]), C.c_int32(len(indptr) - 1), C.c_int32(X.shape[1]))
    return Y


def knn_rank(X, query: int, k: int):
    X = np.ascontiguousarray(X, np.float32)
    idx, dist = np.empty(k, np.int32), np.empty(k, np.float64)
    rc = lib().knn_rank_f64(_p(X), C.c_int64(X.shape[1]), C.c_int32(X.shape[0]), C.c_int32(X.shape[1]), C.c_int32(query), C.c_int32(k),
                            _p(idx), _p(dist))
    assert rc == 0
    return idx, dist


def gae_loss_grad(z, lab_indptr, lab_indices, norm: float, pos_weight: float):
    z = np.ascontiguousarray(z, np.float32)
    n, d = z.shape
    loss = C.c_double(0.0)
    dz = np.empty((n, d), np.float64)
    lp, li = np.ascontiguousarray(lab_indptr, np.int32), np.ascontiguousarray(lab_indices, np.int32)
    rc = lib().gae_loss_grad_f64(_p(z), C.c_int64(d), C.c_int32(n), C.c_int32(d), _p(lp), _p(li), C.c_double(norm), C.c_double(pos_weight),
                                 C.byref(loss), _p(dz))
    assert rc == 0
    return loss.value, dz
