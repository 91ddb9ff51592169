"""`ExtendedKalmanFilter.KalmanFilter` with the reference's names, argument order and return arity
(reference: Python Kalman Filter/ExtendedKalmanFilter.py:5-80), backed by the sm_100a kernels.

Every method accepts what the reference accepts (lists / numpy arrays for one filter) and returns
float64 numpy arrays; arrays may also carry a leading batch dimension N.  For replaying whole logs
use `poseestimationkf_b200.batched.replay`, which runs the loop of main_file.py:38-47 in one launch."""
from __future__ import annotations

import numpy as np

from poseestimationkf_b200 import _lib
from Wahba import Wahba
from _bridge import FixedCall, as_rows, call, from_rows, is_single, is_single_matrix


def _dt_rows(dt_seconds, n):
    dt = np.asarray(dt_seconds, dtype=np.float64)
    return np.broadcast_to(dt.astype(np.float32).reshape(-1), (n,)).reshape(1, n).copy()


class KalmanFilter:
    def __init__(self, T0, mag_0, acc_0, eps):            # reference :6-11 (NB: mag before acc)
        self.previousT = T0
        self.wahba = Wahba(acc_0, mag_0)
        self.Q = np.identity(3)
        self.R = np.identity(4)
        self.eps = eps

    def setQ(self, q):                                     # :12-13 (cumulative, in place)
        self.Q *= q

    def setR(self, r):                                     # :14-15
        self.R *= r

    def Comparator(self, q1, q2):                          # :16-23
        a, batched = as_rows(q1, (4,))
        b, _ = as_rows(q2, (4,))
        lib = _lib.load()
        out, = call([a, b], [4], lambda i, o, n, s: lib.posekf_comparator_f32(n, i[0], i[1], o[0], s))
        return from_rows(out, (4,), batched)

    @staticmethod
    def RungeKutta4(q_0, T, w):                            # :25-41 ; T is a time step in NANOSECONDS
        q, batched = as_rows(q_0, (4,))
        wd, _ = as_rows(w, (3,))
        dt = _dt_rows(np.asarray(T, dtype=np.float64) * (10 ** -9), q.shape[1])
        lib = _lib.load()
        out, = call([q, dt, wd], [4], lambda i, o, n, s: lib.posekf_rk4_f32(n, i[0], i[1], 0, i[2], o[0], s))
        return from_rows(out, (4,), batched)

    def GetJacobian_A(self, w):                            # :43-48
        wd, batched = as_rows(w, (3,))
        lib = _lib.load()
        out, = call([wd], [16], lambda i, o, n, s: lib.posekf_jacobians_f32(n, i[0], o[0], None, None, s))
        return from_rows(out, (4, 4), batched)

    def GetJacobian_B(self, q):                            # :51-56
        qd, batched = as_rows(q, (4,))
        lib = _lib.load()
        out, = call([qd], [12], lambda i, o, n, s: lib.posekf_jacobians_f32(n, None, None, i[0], o[0], s))
        return from_rows(out, (4, 3), batched)

    # fixed layouts of the one-filter calls (see _bridge.FixedCall): inputs / outputs per argument
    _predict1 = FixedCall([(3,), (1,), (4,), (4, 4), (3, 3), (4, 4)], [(4,), (4, 4), (4, 4)])
    _correct1 = FixedCall([(3,), (3,), (3,), (3,), (4,), (4, 4), (4, 4)], [(4,), (4, 4)])

    def Prediction(self, Gyro, T, X_k, P_k):               # :58-68
        if is_single(Gyro, 3) and is_single(X_k, 4) and is_single_matrix(P_k, 4, 4) and np.ndim(T) == 0:
            dt = (float(T) - float(self.previousT)) * (10 ** -9)
            z, pn, k = self._predict1.run(
                [Gyro, dt, X_k, P_k, self.Q, self.R],
                lambda lib, i, o, s: lib.posekf_predict_f32(1, i[0], i[1], 0, i[2], i[3], i[4], i[5], None, None, o[0], o[1], o[2], s))
            self.previousT = T                             # :67
            return z, pn, k
        g, batched = as_rows(Gyro, (3,))
        x, _ = as_rows(X_k, (4,))
        p, _ = as_rows(P_k, (4, 4))
        n = g.shape[1]
        dt = _dt_rows((np.asarray(T, dtype=np.float64) - np.asarray(self.previousT, dtype=np.float64)) * (10 ** -9), n)
        qm = np.ascontiguousarray(self.Q, dtype=np.float32).reshape(-1)
        rm = np.ascontiguousarray(self.R, dtype=np.float32).reshape(-1)
        lib = _lib.load()
        z, pn, k = call([g, dt, x, p, qm, rm], [4, 16, 16],
                        lambda i, o, n_, s: lib.posekf_predict_f32(n_, i[0], i[1], 0, i[2], i[3], i[4], i[5], None, None,
                                                                   o[0], o[1], o[2], s))
        self.previousT = T                                 # :67
        return from_rows(z, (4,), batched), from_rows(pn, (4, 4), batched), from_rows(k, (4, 4), batched)

    def Correction(self, Mag, Acc, z_k, P_k, K_k):         # :70-80 (NB: Mag before Acc)
        w = self.wahba
        if (is_single(Mag, 3) and is_single(Acc, 3) and is_single(z_k, 4) and is_single_matrix(P_k, 4, 4)
                and is_single_matrix(K_k, 4, 4) and is_single(w.w_initial_acc, 3) and is_single(w.w_initial_mag, 3)):
            algo = _lib.WAHBA[w.algo]
            x, pn = self._correct1.run(
                [Mag, Acc, w.w_initial_acc, w.w_initial_mag, z_k, P_k, K_k],
                lambda lib, i, o, s: lib.posekf_correct_f32(1, i[0], i[1], i[2], i[3], 1, i[4], i[5], i[6], o[0], o[1], None, None, algo, s))
            return x, pn
        m, batched = as_rows(Mag, (3,))
        a, _ = as_rows(Acc, (3,))
        z, _ = as_rows(z_k, (4,))
        p, _ = as_rows(P_k, (4, 4))
        k, _ = as_rows(K_k, (4, 4))
        ra, rm, shared = self.wahba._ref_rows(a.shape[1])
        lib = _lib.load()
        algo = _lib.WAHBA[self.wahba.algo]
        x, pn = call([m, a, ra, rm, z, p, k], [4, 16],
                     lambda i, o, n, s: lib.posekf_correct_f32(n, i[0], i[1], i[2], i[3], shared, i[4], i[5], i[6], o[0], o[1],
                                                               None, None, algo, s))
        return from_rows(x, (4,), batched), from_rows(pn, (4, 4), batched)
