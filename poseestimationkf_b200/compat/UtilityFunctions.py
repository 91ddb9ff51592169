"""`UtilityFunctions` with the reference's names (reference: Python Kalman Filter/UtilityFunctions.py)."""
from __future__ import annotations

import numpy as np

from poseestimationkf_b200 import batched as _b
from _bridge import to_dev, to_host


def Quart2RPY(q):                                          # reference :3-14 (degrees, asin not clamped)
    qd, batched = to_dev(q, (4,))
    return to_host(_b.quat2rpy(qd), (3,), batched)


def norm(a):                                               # :16-21
    arr = np.asarray(a, dtype=np.float64)
    v, _ = to_dev(arr, (arr.shape[-1],))
    out = _b.norm(v).detach().cpu().numpy().astype(np.float64)
    return out if arr.ndim == 2 else out[0]


def DimensionalSplit(S):                                   # :24-34 -- pure re-indexing, no arithmetic
    return [[S[j][i] for j in range(len(S))] for i in range(len(S[0]))]
