"""Mirror of ``dance/utils/matrix.py`` for the functions on the hot path: ``normalize`` (:8-67) and ``pairwise_distance``
(:164-180), executed by the device kernels."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def normalize(mat, *, mode: str = "normalize", axis: int = 0, eps: float = -1.0):
    """Same contract as the reference: numpy in → numpy out, torch in → torch (CUDA) out; 2-D fp32."""
    if isinstance(mat, torch.Tensor):
        is_torch = True
    elif not isinstance(mat, np.ndarray):
        raise TypeError(f"Invalid type for input matrix: {type(mat)}")
    else:
        is_torch = False
    if mode not in ops.NORM_MODE:       # the reference silently returns mat / 1 for unknown modes (denom = None → 1)
        if not (eps == -1 or eps > 0):
            raise ValueError(f"Invalid {eps=!r}. Must be positive or -1, the later set zero entries to one.")
        return mat / 1
    X = mat if is_torch else torch.as_tensor(np.ascontiguousarray(mat, dtype=np.float32))
    X = X.to(device="cuda", dtype=torch.float32).contiguous()
    if X.dim() != 2:
        raise ValueError("normalize: 2-D input expected")
    out = ops.matrix_normalize(X, mode, axis % 2, eps)
    return out if is_torch else out.cpu().numpy()


def pairwise_distance(x: np.ndarray, dist_func_id: int = 0) -> np.ndarray:
    if dist_func_id != 0:
        raise NotImplementedError("only the euclidean distance (dist_func_id=0) is built")
    X = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    return ops.pairwise_l2_dense(X).cpu().numpy()
