"""`ExtendedKalmanFilter.KalmanFilter` with the reference's names, argument order and return arity
(reference: Python Kalman Filter/ExtendedKalmanFilter.py:5-80), backed by the sm_100a kernels.

Every method accepts what the reference accepts (lists / numpy arrays for one filter) and returns
float64 numpy arrays; arrays may also carry a leading batch dimension N.  For replaying whole logs
use `poseestimationkf_b200.batched.replay`, which runs the loop of main_file.py:38-47 in one launch."""
from __future__ import annotations

import numpy as np
import torch

from poseestimationkf_b200 import batched as _b
from Wahba import Wahba
from _bridge import device, to_dev, to_host


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
        a, batched = to_dev(q1, (4,))
        b, _ = to_dev(q2, (4,))
        return to_host(_b.comparator(a, b), (4,), batched)

    @staticmethod
    def RungeKutta4(q_0, T, w):                            # :25-41 ; T is a time step in NANOSECONDS
        q, batched = to_dev(q_0, (4,))
        wd, _ = to_dev(w, (3,))
        dt = np.asarray(T, dtype=np.float64) * (10 ** -9)
        if dt.ndim == 0:
            out = _b.rk4(q, float(dt), wd)
        else:
            out = _b.rk4(q, torch.from_numpy(dt.astype(np.float32)).to(device()), wd)
        return to_host(out, (4,), batched)

    def GetJacobian_A(self, w):                            # :43-48
        wd, batched = to_dev(w, (3,))
        return to_host(_b.jacobian_a(wd), (4, 4), batched)

    def GetJacobian_B(self, q):                            # :51-56
        qd, batched = to_dev(q, (4,))
        return to_host(_b.jacobian_b(qd), (4, 3), batched)

    def Prediction(self, Gyro, T, X_k, P_k):               # :58-68
        g, batched = to_dev(Gyro, (3,))
        x, _ = to_dev(X_k, (4,))
        p, _ = to_dev(P_k, (4, 4))
        dt = (np.asarray(T, dtype=np.float64) - np.asarray(self.previousT, dtype=np.float64)) * (10 ** -9)
        dt_arg = float(dt) if dt.ndim == 0 else torch.from_numpy(dt.astype(np.float32)).to(device())
        dev = device()
        qm = torch.from_numpy(np.ascontiguousarray(self.Q, dtype=np.float32).reshape(-1)).to(dev)
        rm = torch.from_numpy(np.ascontiguousarray(self.R, dtype=np.float32).reshape(-1)).to(dev)
        z, pn, k = _b.predict(g, dt_arg, x, p, qm, rm)
        self.previousT = T                                 # :67
        return to_host(z, (4,), batched), to_host(pn, (4, 4), batched), to_host(k, (4, 4), batched)

    def Correction(self, Mag, Acc, z_k, P_k, K_k):         # :70-80 (NB: Mag before Acc)
        m, batched = to_dev(Mag, (3,))
        a, _ = to_dev(Acc, (3,))
        z, _ = to_dev(z_k, (4,))
        p, _ = to_dev(P_k, (4, 4))
        k, _ = to_dev(K_k, (4, 4))
        ra, rm = self.wahba._refs(a.shape[1])
        x, pn, _, _ = _b.correct(m, a, ra, rm, z, p, k)
        return to_host(x, (4,), batched), to_host(pn, (4, 4), batched)
