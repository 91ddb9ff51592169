"""`Wahba.Wahba` with the reference's interface (reference: Python Kalman Filter/Wahba.py:3-50),
backed by the sm_100a kernels.  `algo` selects the device algorithm ("qr2": rank-2 SVD with a closed-form 2x2 polar
factor; "jacobi": QR-preconditioned one-sided Jacobi SVD -- both accurate to float32 rounding for any weights)."""
from __future__ import annotations

import numpy as np

from poseestimationkf_b200 import _lib
from _bridge import FixedCall, as_rows, call, from_rows, is_single


class Wahba:
    algo = "qr2"

    def __init__(self, acc, mag):                          # reference :4-6
        self.w_initial_acc = acc
        self.w_initial_mag = mag

    def _ref_rows(self, n):
        """reference vectors as rows; one shared pair ([3] each) or one per filter ([3, n])"""
        ra, ba = as_rows(self.w_initial_acc, (3,))
        rm, _ = as_rows(self.w_initial_mag, (3,))
        if ba:
            if ra.shape[1] != n:
                raise ValueError("batched reference vectors must match the batch size")
            return ra, rm, 0
        return ra.reshape(3), rm.reshape(3), 1

    @staticmethod
    def _weights(k, n):
        k = np.asarray(k, dtype=np.float64)
        return np.broadcast_to(k.astype(np.float32).reshape(-1), (n,)).reshape(1, n).copy()

    _solve1 = FixedCall([(3,), (3,), (3,), (3,)], [(3, 3), (4,)])

    def _solve(self, acc, mag, k_acc, k_mag, want_rotation):
        if (is_single(acc, 3) and is_single(mag, 3) and is_single(self.w_initial_acc, 3) and is_single(self.w_initial_mag, 3)
                and np.ndim(k_acc) == 0 and np.ndim(k_mag) == 0):
            algo, ka, km = _lib.WAHBA[self.algo], float(k_acc), float(k_mag)
            rot, quat = self._solve1.run(
                [self.w_initial_acc, self.w_initial_mag, acc, mag],
                lambda lib, i, o, s: lib.posekf_wahba_f32(1, i[0], i[1], 1, i[2], i[3], None, None, ka, km, 0,
                                                          o[0] if want_rotation else None, None if want_rotation else o[1], algo, 0, s))
            return rot if want_rotation else quat
        a, batched = as_rows(acc, (3,))
        m, _ = as_rows(mag, (3,))
        n = a.shape[1]
        ra, rm, shared = self._ref_rows(n)
        lib = _lib.load()
        algo = _lib.WAHBA[self.algo]
        if want_rotation:
            out, = call([ra, rm, a, m, self._weights(k_acc, n), self._weights(k_mag, n)], [9],
                        lambda i, o, n_, s: lib.posekf_wahba_f32(n_, i[0], i[1], shared, i[2], i[3], i[4], i[5], 0.0, 0.0, 0,
                                                                 o[0], None, algo, 0, s))
            return from_rows(out, (3, 3), batched)
        out, = call([ra, rm, a, m, self._weights(k_acc, n), self._weights(k_mag, n)], [4],
                    lambda i, o, n_, s: lib.posekf_wahba_f32(n_, i[0], i[1], shared, i[2], i[3], i[4], i[5], 0.0, 0.0, 0,
                                                             None, o[0], algo, 0, s))
        return from_rows(out, (4,), batched)

    def getRotation(self, acc, mag, k_acc, k_mag):         # :8-17
        return self._solve(acc, mag, k_acc, k_mag, True)

    @staticmethod
    def RotationMatrix2Quart(M):                           # :20-47
        r, batched = as_rows(M, (3, 3))
        lib = _lib.load()
        out, = call([r], [4], lambda i, o, n, s: lib.posekf_rot2quat_f32(n, i[0], o[0], s))
        return from_rows(out, (4,), batched)

    def getQuarternion(self, acc, mag, k_acc, k_mag):      # :49-50
        return self._solve(acc, mag, k_acc, k_mag, False)
