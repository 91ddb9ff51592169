"""CPU oracle for the batched quaternion EKF hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a float64 numpy restatement of the reference's offline replay path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
it; the product package (`poseestimationkf_b200/`) never does and fails loudly without its CUDA
library.

Parity status: PINNED.  `tests/test_oracle.py` checks this file against
  * the reference's two own known answers: the SVD-Wahba hand check
    (`Python Kalman Filter/WahbaProblem_singularValue.py:4-26`) and the RK4 convergence run
    (`Quarternions.py:44-60,99-112`);
  * outputs of the unmodified reference classes, generated in the build container by
    `tests/golden/make_golden.py` (which imports `/root/reference/Python Kalman Filter`) and
    committed under `tests/golden/*.npz`.

Two forms are provided:
  * the *scalar* form (`OracleWahba`, `OracleEKF`, `replay_scalar`): one filter, one numpy call per
    reference numpy call, same operation order, so that its wall-clock per step is the reference's
    and it can serve as the timed CPU baseline ("port");
  * the *batched* form (`replay_batched`, `wahba_batched`): the same arithmetic vectorised over a
    leading filter axis, used as the fast checker for thousands of filters.  It is pinned to the
    scalar form by tests.

File:line citations are relative to /root/reference/.
"""
from __future__ import annotations

import math

import numpy as np

PKF = "Python Kalman Filter"  # directory shorthand used in citations


# --------------------------------------------------------------------------------------------
# UtilityFunctions.py
# --------------------------------------------------------------------------------------------
def norm(a):
    """L2 norm, accumulated left to right.  (PKF/UtilityFunctions.py:16-21)"""
    acc = 0.0
    for v in a:
        acc += v ** 2
    return np.sqrt(acc)


def quat_to_rpy_deg(q):
    """Roll/pitch/yaw in degrees, asin NOT clamped.  (PKF/UtilityFunctions.py:3-14)"""
    out = np.zeros(3)
    out[0] = np.arctan2(2 * (q[0] * q[1] + q[2] * q[3]), 1 - 2 * (q[1] * q[1] + q[2] * q[2]))
    out[1] = np.arcsin(2 * (q[0] * q[2] - q[3] * q[1]))
    out[2] = np.arctan2(2 * (q[0] * q[3] + q[1] * q[2]), 1 - 2 * (q[2] * q[2] + q[3] * q[3]))
    return out * 180.0 / np.pi


def dimensional_split(rows):
    """list-of-rows -> list-of-columns.  (PKF/UtilityFunctions.py:24-34)"""
    return [[rows[j][i] for j in range(len(rows))] for i in range(len(rows[0]))]


# --------------------------------------------------------------------------------------------
# Small matrix builders shared by RK4 and the Jacobians
# --------------------------------------------------------------------------------------------
def half_omega(w):
    """0.5 * Omega(w): the 4x4 matrix of q_dot = 0.5*Omega(w) q.
    (PKF/ExtendedKalmanFilter.py:27-30 and :44-47 -- the same literal appears twice.)"""
    wx, wy, wz = w[0], w[1], w[2]
    return 0.5 * np.asarray([[0.0, -wx, -wy, -wz],
                             [wx, 0.0, wz, -wy],
                             [wy, -wz, 0.0, wx],
                             [wz, wy, -wx, 0.0]])


def jacobian_b(q):
    """4x3 noise Jacobian.  (PKF/ExtendedKalmanFilter.py:51-56)"""
    return 0.5 * np.asarray([[-q[1], -q[2], -q[3]],
                             [q[0], q[3], -q[2]],
                             [-q[3], q[0], q[1]],
                             [q[2], -q[1], q[0]]])


def rk4(q0, dt_ns, w):
    """One RK4 step of q_dot = 0.5*Omega(w) q over dt_ns nanoseconds, then normalise.
    (PKF/ExtendedKalmanFilter.py:25-41; `T*(10**-9)` at :32 is kept as written.)"""
    W = half_omega(w)
    h = dt_ns * (10 ** -9)
    k1 = np.dot(W, q0)
    k2 = np.dot(W, q0 + h / 2 * k1)
    k3 = np.dot(W, q0 + h / 2 * k2)
    k4 = np.dot(W, q0 + h * k3)
    q1 = q0 + 1 / 6 * h * (k1 + 2 * k2 + 2 * k3 + k4)
    return q1 / norm(q1)


def rk4_seconds(q0, h, w):
    """The variant in /root/reference/Quarternions.py:44-60 (step in seconds) -- used only to check
    the reference's RK4 convergence known answer."""
    W = half_omega(w)
    k1 = np.dot(W, q0)
    k2 = np.dot(W, q0 + h / 2 * k1)
    k3 = np.dot(W, q0 + h / 2 * k2)
    k4 = np.dot(W, q0 + h * k3)
    q1 = q0 + 1 / 6 * h * (k1 + 2 * k2 + 2 * k3 + k4)
    return q1 / norm(q1)


# --------------------------------------------------------------------------------------------
# Wahba.py
# --------------------------------------------------------------------------------------------
def rotation_to_quat(M):
    """3-branch rotation-matrix -> quaternion [w,x,y,z]; strict '>' comparisons, ties fall to the
    last branch; no trace-positive branch; output not renormalised; NaN at M == I.
    (PKF/Wahba.py:20-47)"""
    t1 = 1.0 + M[0][0] - M[1][1] - M[2][2]
    t2 = 1.0 - M[0][0] + M[1][1] - M[2][2]
    t3 = 1.0 - M[0][0] - M[1][1] + M[2][2]
    if (t1 > t2) and (t1 > t3):
        S = np.sqrt(t1) * 2
        q = [(M[2][1] - M[1][2]) / S, 0.25 * S, (M[0][1] + M[1][0]) / S, (M[0][2] + M[2][0]) / S]
    elif (t2 > t1) and (t2 > t3):
        S = np.sqrt(t2) * 2
        q = [(M[0][2] - M[2][0]) / S, (M[0][1] + M[1][0]) / S, 0.25 * S, (M[1][2] + M[2][1]) / S]
    else:
        S = np.sqrt(t3) * 2
        q = [(M[1][0] - M[0][1]) / S, (M[0][2] + M[2][0]) / S, (M[1][2] + M[2][1]) / S, 0.25 * S]
    return np.asarray(q)


class OracleWahba:
    """Two-observation Wahba solver by SVD.  (PKF/Wahba.py:3-50)"""

    def __init__(self, acc_ref, mag_ref):
        self.acc_ref = acc_ref      # PKF/Wahba.py:5  w_initial_acc
        self.mag_ref = mag_ref      # PKF/Wahba.py:6  w_initial_mag

    def rotation(self, acc, mag, k_acc, k_mag):
        """B = k_acc*outer(acc_ref, acc) + k_mag*outer(mag_ref, mag); R = U diag(1,1,det U det Vt) Vt.
        (PKF/Wahba.py:8-17)"""
        B = k_acc * np.outer(self.acc_ref, acc) + k_mag * np.outer(self.mag_ref, mag)
        u, _s, vh = np.linalg.svd(B)
        M = np.diag(np.asarray([1, 1, np.linalg.det(u) * np.linalg.det(vh)]))
        return np.matmul(np.matmul(u, M), vh)

    def quaternion(self, acc, mag, k_acc, k_mag):
        """(PKF/Wahba.py:49-50)"""
        return rotation_to_quat(self.rotation(acc, mag, k_acc, k_mag))


# --------------------------------------------------------------------------------------------
# ExtendedKalmanFilter.py
# --------------------------------------------------------------------------------------------
def comparator(q1, q2):
    """conj(q1) (x) q2 as a 4x4 mat-vec; element [0] equals dot(q1, q2).
    (PKF/ExtendedKalmanFilter.py:16-23)"""
    c = np.asarray([q1[0], -q1[1], -q1[2], -q1[3]])
    L = np.asarray([[c[0], -c[1], -c[2], -c[3]],
                    [c[1], c[0], -c[3], c[2]],
                    [c[2], c[3], c[0], -c[1]],
                    [c[3], -c[2], c[1], c[0]]])
    return np.matmul(L, np.asarray([q2[0], q2[1], q2[2], q2[3]]))


class OracleEKF:
    """The reference's `KalmanFilter`.  NB constructor order is (T0, mag_0, acc_0, eps) while the
    Wahba solver takes (acc, mag).  (PKF/ExtendedKalmanFilter.py:5-80)"""

    def __init__(self, t0_ns, mag_ref, acc_ref, eps=0.5):
        self.prev_t = t0_ns                           # :7
        self.wahba = OracleWahba(acc_ref, mag_ref)    # :8
        self.Q = np.identity(3)                       # :9
        self.R = np.identity(4)                       # :10
        self.eps = eps                                # :11 (never read)

    def set_q(self, q):
        self.Q *= q                                   # :12-13  cumulative, in place

    def set_r(self, r):
        self.R *= r                                   # :14-15

    def predict(self, gyro, t_ns, x, P):
        """(PKF/ExtendedKalmanFilter.py:58-68)  A = 0.5*Omega(gyro) is used as the transition matrix
        as-is (continuous-time Jacobian; reference behaviour); B is evaluated
        at the pre-propagation state."""
        A = half_omega(gyro)
        Bn = jacobian_b(x)
        P = np.matmul(np.matmul(A, P), A.transpose()) + np.matmul(np.matmul(Bn, self.Q), Bn.transpose())
        z = rk4(x, t_ns - self.prev_t, gyro)
        S = P + self.R
        K = np.matmul(P, np.linalg.inv(S))
        self.prev_t = t_ns
        return z, P, K

    def correct(self, mag, acc, z, P, K):
        """(PKF/ExtendedKalmanFilter.py:70-80)  NB argument order (Mag, Acc).  Returns (X, P); the
        q/-q decision is also returned by `correct_ex`."""
        X, P, _flip, _y = self.correct_ex(mag, acc, z, P, K)
        return X, P

    def correct_ex(self, mag, acc, z, P, K):
        y = self.wahba.quaternion(acc, mag, abs(acc[2]), 1 - (abs(acc[2])))   # :71
        flip = bool(comparator(y, z)[0] < 0.0)                                # :73-74
        if flip:
            y = -y                                                            # :75
        X = z + np.matmul(K, y - z)                                           # :76-77
        P = P - np.matmul(K, P)                                               # :78
        X = X / norm(X)                                                       # :79
        return X, P, flip, y


def replay_scalar(t_ns, gyro, acc, mag, acc_ref, mag_ref, q_scale=1.0, r_scale=0.1,
                  x0=None, P0=None, return_aux=False):
    """The caller contract of PKF/main_file.py:19-47 for ONE filter.

    t_ns : [T+1] integer timestamps; t_ns[0] seeds previousT and is dropped (:19,:25)
    gyro, acc, mag : [T,3]; acc_ref/mag_ref : [3] (the log's acc_0 / mag_0)
    Returns X [T,4] (state after each step; the initial state is not included), final P [4,4] and,
    if return_aux, the per-step flip mask [T] and the sign-fixed Wahba quaternion [T,4].
    """
    ekf = OracleEKF(t_ns[0], mag_ref, acc_ref, 0.5)        # main_file.py:19
    ekf.set_q(q_scale)                                     # :21
    ekf.set_r(r_scale)                                     # :22
    P = np.identity(4) if P0 is None else np.array(P0, dtype=np.float64)          # :23
    X = np.asarray([1.0, 0.0, 0.0, 0.0]) if x0 is None else np.array(x0, dtype=np.float64)  # :26
    T = len(gyro)
    traj = np.empty((T, 4))
    flips = np.zeros(T, dtype=bool)
    ys = np.empty((T, 4))
    for i in range(T):                                     # :38
        z, P, K = ekf.predict(gyro[i], t_ns[i + 1], X, P)  # :39
        X, P, flips[i], ys[i] = ekf.correct_ex(mag[i], acc[i], z, P, K)   # :43
        traj[i] = X                                        # :44
    if return_aux:
        return traj, P, flips, ys
    return traj, P


def lowpass_scalar(x, alpha, y0=None):
    """y <- alpha*x + (1-alpha)*y, state starts at 0.  (PKF/Test.py:27-33; C++ twin
    `Kalman Filter Server/PoseEstimator/KalmanFilter.cpp:21-24`, initial state :16-18.)
    x : [T,3] -> [T,3]"""
    y = np.zeros(3) if y0 is None else np.array(y0, dtype=np.float64)
    out = np.empty_like(np.asarray(x, dtype=np.float64))
    for i in range(len(x)):
        y = alpha * np.asarray(x[i], dtype=np.float64) + (1 - alpha) * y
        out[i] = y
    return out


def interpolate_sensor(y1, y2, t1, t2, t3):
    """Linear interpolation of a sensor sample pair to the gyro timestamp, float64, in the reference's order of
    operations: (y2 - y1) / (t2 - t1) * (t3 - t1) + y1 with the int64 nanosecond stamps converted to double first.
    (`Kalman Filter Server/PoseEstimator/Parser.cpp:259-267` LinearInterpolationSensor.)
    y1, y2: [...,3]; t1, t2, t3: integer ns (broadcastable)."""
    y1 = np.asarray(y1, dtype=np.float64)
    y2 = np.asarray(y2, dtype=np.float64)
    t1, t2, t3 = (np.asarray(t, dtype=np.float64)[..., None] for t in (t1, t2, t3))
    return (y2 - y1) / (t2 - t1) * (t3 - t1) + y1


def normalize_values(v):
    """`Parser::NormalizeValues` (`Parser.cpp:221-228`): divide by sqrt(x*x + y*y + z*z), summed left to right."""
    v = np.asarray(v, dtype=np.float64)
    den = np.sqrt((v[..., 0] * v[..., 0]) + (v[..., 1] * v[..., 1]) + (v[..., 2] * v[..., 2]))
    return v / den[..., None]


def interpolate_normalise(y1, y2, t1, t2, t3):
    """ExecuteKalmanFilter's sequence for one sensor (`Parser.cpp:232-242`): interpolate to the gyro timestamp, then
    normalise.  PINNED: tests/test_oracle.py holds it bit for bit to tests/golden/preprocess_ref.npz, which
    tests/golden/make_golden_cpp.py produced by executing the reference's own C++ text (oracle/_ref)."""
    return normalize_values(interpolate_sensor(y1, y2, t1, t2, t3))


def initial_values(samples):
    """`InitialValues` (`Kalman Filter Server/PoseEstimator/InitialValues.cpp:19-66`): running sum of the K samples
    in arrival order, mean = sum / K, unbiased variance = sum_i (x_i - mean)^2 / (K - 1), accumulated in order.
    samples [..., K, 3] -> (mean [...,3], var [...,3]).  PINNED bit for bit to tests/golden/initial_values_ref.npz
    (the unmodified InitialValues.cpp compiled into oracle/_ref)."""
    x = np.asarray(samples, dtype=np.float64)
    K = x.shape[-2]
    acc = np.zeros(x.shape[:-2] + (3,))
    for i in range(K):
        acc = acc + x[..., i, :]
    mean = acc / K
    var = np.zeros_like(mean)
    for i in range(K):
        var = var + (x[..., i, :] - mean) * (x[..., i, :] - mean)
    return mean, var / float(K - 1)


# --------------------------------------------------------------------------------------------
# Batched (vectorised over filters) float64 form: the fast checker
# --------------------------------------------------------------------------------------------
def _half_omega_b(w):
    """[N,3] -> [N,4,4]"""
    N = w.shape[0]
    A = np.zeros((N, 4, 4))
    wx, wy, wz = 0.5 * w[:, 0], 0.5 * w[:, 1], 0.5 * w[:, 2]
    A[:, 0, 1], A[:, 0, 2], A[:, 0, 3] = -wx, -wy, -wz
    A[:, 1, 0], A[:, 1, 2], A[:, 1, 3] = wx, wz, -wy
    A[:, 2, 0], A[:, 2, 1], A[:, 2, 3] = wy, -wz, wx
    A[:, 3, 0], A[:, 3, 1], A[:, 3, 2] = wz, wy, -wx
    return A


def _jacobian_b_b(q):
    N = q.shape[0]
    B = np.empty((N, 4, 3))
    q0, q1, q2, q3 = (0.5 * q[:, i] for i in range(4))
    B[:, 0, 0], B[:, 0, 1], B[:, 0, 2] = -q1, -q2, -q3
    B[:, 1, 0], B[:, 1, 1], B[:, 1, 2] = q0, q3, -q2
    B[:, 2, 0], B[:, 2, 1], B[:, 2, 2] = -q3, q0, q1
    B[:, 3, 0], B[:, 3, 1], B[:, 3, 2] = q2, -q1, q0
    return B


def _norm_b(v):
    acc = np.zeros(v.shape[0])
    for i in range(v.shape[1]):
        acc = acc + v[:, i] ** 2
    return np.sqrt(acc)


def rotation_to_quat_batched(M):
    """[N,3,3] -> [N,4], same branch rules as `rotation_to_quat`."""
    t1 = 1.0 + M[:, 0, 0] - M[:, 1, 1] - M[:, 2, 2]
    t2 = 1.0 - M[:, 0, 0] + M[:, 1, 1] - M[:, 2, 2]
    t3 = 1.0 - M[:, 0, 0] - M[:, 1, 1] + M[:, 2, 2]
    b1 = (t1 > t2) & (t1 > t3)
    b2 = ~b1 & (t2 > t1) & (t2 > t3)
    b3 = ~(b1 | b2)
    q = np.empty((M.shape[0], 4))
    with np.errstate(invalid="ignore", divide="ignore"):
        for mask, t, w_num, comps in (
            (b1, t1, (2, 1, 1, 2), ((1, None), (2, (0, 1, 1, 0)), (3, (0, 2, 2, 0)))),
            (b2, t2, (0, 2, 2, 0), ((2, None), (1, (0, 1, 1, 0)), (3, (1, 2, 2, 1)))),
            (b3, t3, (1, 0, 0, 1), ((3, None), (1, (0, 2, 2, 0)), (2, (1, 2, 2, 1)))),
        ):
            if not mask.any():
                continue
            Mm = M[mask]
            S = np.sqrt(t[mask]) * 2
            qq = np.empty((Mm.shape[0], 4))
            a, b, c, d = w_num
            qq[:, 0] = (Mm[:, a, b] - Mm[:, c, d]) / S
            for idx, spec in comps:
                if spec is None:
                    qq[:, idx] = 0.25 * S
                else:
                    a, b, c, d = spec
                    qq[:, idx] = (Mm[:, a, b] + Mm[:, c, d]) / S
            q[mask] = qq
    return q


def wahba_rotation_batched(acc_ref, mag_ref, acc, mag, k_acc, k_mag):
    """[N,3] x4, [N] x2 -> R [N,3,3]  (PKF/Wahba.py:8-17 vectorised)."""
    B = (k_acc[:, None, None] * acc_ref[:, :, None] * acc[:, None, :]
         + k_mag[:, None, None] * mag_ref[:, :, None] * mag[:, None, :])
    u, _s, vh = np.linalg.svd(B)
    d = np.linalg.det(u) * np.linalg.det(vh)
    u = u.copy()
    u[:, :, 2] *= d[:, None]
    return np.matmul(u, vh)


def wahba_batched(acc_ref, mag_ref, acc, mag, k_acc, k_mag):
    return rotation_to_quat_batched(wahba_rotation_batched(acc_ref, mag_ref, acc, mag, k_acc, k_mag))


def replay_batched(dt_ns, gyro, acc, mag, acc_ref, mag_ref, q_scale, r_scale, x0=None, P0=None,
                   store=True):
    """Vectorised `replay_scalar` for N filters.

    dt_ns : [T] (shared) or [T,N] step lengths in ns (already differenced);
    gyro/acc/mag : [T,3,N] (the product's stream layout); acc_ref/mag_ref : [N,3];
    q_scale/r_scale : scalar or [N].
    Returns dict(X=[T,N,4] if store else None, X_final [N,4], P_final [N,4,4], flips [T,N], y [T,N,4]).
    """
    T, _, N = gyro.shape
    qs = np.broadcast_to(np.asarray(q_scale, dtype=np.float64), (N,))
    rs = np.broadcast_to(np.asarray(r_scale, dtype=np.float64), (N,))
    Q = qs[:, None, None] * np.identity(3)[None]
    R = rs[:, None, None] * np.identity(4)[None]
    X = np.tile(np.asarray([1.0, 0.0, 0.0, 0.0]), (N, 1)) if x0 is None else np.array(x0, dtype=np.float64)
    P = np.tile(np.identity(4), (N, 1, 1)) if P0 is None else np.array(P0, dtype=np.float64)
    acc_ref = np.asarray(acc_ref, dtype=np.float64)
    mag_ref = np.asarray(mag_ref, dtype=np.float64)
    traj = np.empty((T, N, 4)) if store else None
    flips = np.zeros((T, N), dtype=bool)
    ys = np.empty((T, N, 4)) if store else None
    dt_ns = np.asarray(dt_ns, dtype=np.float64)
    for i in range(T):
        w = np.ascontiguousarray(gyro[i].T, dtype=np.float64)
        a = np.ascontiguousarray(acc[i].T, dtype=np.float64)
        m = np.ascontiguousarray(mag[i].T, dtype=np.float64)
        A = _half_omega_b(w)
        Bn = _jacobian_b_b(X)
        P = A @ P @ A.transpose(0, 2, 1) + Bn @ Q @ Bn.transpose(0, 2, 1)
        h = (dt_ns[i] * (10 ** -9))
        h = h[:, None] if np.ndim(h) else h
        k1 = np.einsum("nij,nj->ni", A, X)
        k2 = np.einsum("nij,nj->ni", A, X + h / 2 * k1)
        k3 = np.einsum("nij,nj->ni", A, X + h / 2 * k2)
        k4 = np.einsum("nij,nj->ni", A, X + h * k3)
        z = X + 1 / 6 * h * (k1 + 2 * k2 + 2 * k3 + k4)
        z = z / _norm_b(z)[:, None]
        S = P + R
        K = P @ np.linalg.inv(S)
        ka = np.abs(a[:, 2])
        y = wahba_batched(acc_ref, mag_ref, a, m, ka, 1 - ka)
        flip = np.einsum("ni,ni->n", y, z) < 0.0
        # comparator()[0] is dot(y,z) evaluated as y0*z0 + y1*z1 + y2*z2 + y3*z3 by the mat-vec
        y = np.where(flip[:, None], -y, y)
        X = z + np.einsum("nij,nj->ni", K, y - z)
        P = P - K @ P
        X = X / _norm_b(X)[:, None]
        flips[i] = flip
        if store:
            traj[i] = X
            ys[i] = y
    return dict(X=traj, X_final=X, P_final=P, flips=flips, y=ys)


def quat_angle(qa, qb):
    """Quaternion angle error (rad), sign-insensitive: 2*asin(|qa - s*qb|/2), s = sign(qa.qb)
    (the definition in BASELINE.md section 4.4).  This is the angle between the two unit
    4-vectors; the physical rotation between the two attitudes is twice that (`rotation_angle`).
    Inputs [...,4]."""
    qa = np.asarray(qa, dtype=np.float64)
    qb = np.asarray(qb, dtype=np.float64)
    qa = qa / np.linalg.norm(qa, axis=-1, keepdims=True)
    qb = qb / np.linalg.norm(qb, axis=-1, keepdims=True)
    s = np.where(np.sum(qa * qb, axis=-1, keepdims=True) < 0, -1.0, 1.0)
    d = np.linalg.norm(qa - s * qb, axis=-1)
    return 2.0 * np.arcsin(np.clip(d / 2.0, 0.0, 1.0))


def rotation_angle(qa, qb):
    """Physical rotation angle (rad) between the attitudes qa and qb = 2 * quat_angle."""
    return 2.0 * quat_angle(qa, qb)
