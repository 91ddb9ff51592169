"""`UtilityFunctions` with the reference's names (reference: Python Kalman Filter/UtilityFunctions.py)."""
from __future__ import annotations

import numpy as np

from poseestimationkf_b200 import _lib
from _bridge import as_rows, call, from_rows


def Quart2RPY(q):                                          # reference :3-14 (degrees, asin not clamped)
    qd, batched = as_rows(q, (4,))
    lib = _lib.load()
    out, = call([qd], [3], lambda i, o, n, s: lib.posekf_quat2rpy_f32(n, i[0], o[0], s))
    return from_rows(out, (3,), batched)


def norm(a):                                               # :16-21
    arr = np.asarray(a, dtype=np.float64)
    k = arr.shape[-1]
    v, _ = as_rows(arr, (k,))
    lib = _lib.load()
    out, = call([v], [1], lambda i, o, n, s: lib.posekf_norm_f32(n, k, i[0], o[0], s))
    res = out.astype(np.float64).reshape(-1)
    return res if arr.ndim == 2 else res[0]


def DimensionalSplit(S):                                   # :24-34 -- pure re-indexing, no arithmetic
    return [[S[j][i] for j in range(len(S))] for i in range(len(S[0]))]
