"""`Wahba.Wahba` with the reference's interface (reference: Python Kalman Filter/Wahba.py:3-50),
backed by the sm_100a kernels.  `algo` selects the device algorithm ("qr2": rank-2 SVD, accurate
for any weights in float32; "jacobi": B formed as the reference forms it + one-sided Jacobi SVD)."""
from __future__ import annotations

import numpy as np
import torch

from poseestimationkf_b200 import batched as _b
from _bridge import to_dev, to_host


class Wahba:
    algo = "qr2"

    def __init__(self, acc, mag):                          # reference :4-6
        self.w_initial_acc = acc
        self.w_initial_mag = mag

    def _refs(self, n):
        ra, ba = to_dev(self.w_initial_acc, (3,))
        rm, _ = to_dev(self.w_initial_mag, (3,))
        if not ba and n > 1:
            ra, rm = ra.expand(3, n).contiguous(), rm.expand(3, n).contiguous()
        return ra, rm

    def _weights(self, k, n, dev):
        k = np.asarray(k, dtype=np.float64)
        if k.ndim == 0:
            return torch.full((n,), float(k), dtype=torch.float32, device=dev)
        return torch.from_numpy(k.astype(np.float32)).to(dev)

    def _solve(self, acc, mag, k_acc, k_mag, want_rotation):
        a, batched = to_dev(acc, (3,))
        m, _ = to_dev(mag, (3,))
        n = a.shape[1]
        ra, rm = self._refs(n)
        R, q = _b.wahba(ra, rm, a, m, k_acc=self._weights(k_acc, n, a.device), k_mag=self._weights(k_mag, n, a.device),
                        want_rotation=want_rotation, want_quaternion=not want_rotation, algo=self.algo)
        return (to_host(R, (3, 3), batched) if want_rotation else to_host(q, (4,), batched))

    def getRotation(self, acc, mag, k_acc, k_mag):         # :8-17
        return self._solve(acc, mag, k_acc, k_mag, True)

    @staticmethod
    def RotationMatrix2Quart(M):                           # :20-47
        r, batched = to_dev(M, (3, 3))
        return to_host(_b.rot2quat(r), (4,), batched)

    def getQuarternion(self, acc, mag, k_acc, k_mag):      # :49-50
        return self._solve(acc, mag, k_acc, k_mag, False)
