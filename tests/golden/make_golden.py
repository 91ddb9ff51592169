"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; numpy only)

The reference cannot travel to the GPU box, so its outputs are frozen here as small fixtures; the
oracle (oracle/ekf_oracle.py) and the CUDA path are both tested against them.  Inputs are rounded to
float32 BEFORE they are fed to the float64 reference, so input quantisation is not counted as error.
Nothing from the reference is copied into the repository: the modules are imported from where they
lie, and `Quarternions.py`'s RK4 function is exec'd from its own source text (the script itself
imports matplotlib, which is absent).
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("POSEKF_REF", "/root/reference")
PKF = os.path.join(REF, "Python Kalman Filter")
sys.path.insert(0, PKF)
sys.path.insert(0, ROOT)

from ExtendedKalmanFilter import KalmanFilter   # noqa: E402  (the reference)
from Wahba import Wahba                         # noqa: E402
from UtilityFunctions import norm, Quart2RPY, DimensionalSplit   # noqa: E402

from poseestimationkf_b200.synth import make_imu   # noqa: E402


def drive(t_ns, gyro, acc, mag, acc0, mag0, q, r, X0=None, P0=None):
    """The loop of Python Kalman Filter/main_file.py:19-47 around the reference classes (X0 / P0: a caller's own
    initial state instead of main_file.py:23,26)."""
    k = KalmanFilter(t_ns[0], mag0, acc0, 0.5)
    k.setQ(q)
    k.setR(r)
    P = np.identity(4) if P0 is None else np.array(P0, dtype=np.float64)
    X = np.asarray([1., 0., 0., 0.]) if X0 is None else np.array(X0, dtype=np.float64)
    Xs, ys, flips = [], [], []
    for i in range(len(gyro)):
        z, P, K = k.Prediction(gyro[i], t_ns[i + 1], X, P)
        # the measurement quaternion and the q/-q decision, recomputed exactly as Correction does
        y = k.wahba.getQuarternion(acc[i], mag[i], abs(acc[i][2]), 1 - abs(acc[i][2]))
        flip = k.Comparator(y, z)[0] < 0.0
        X, P = k.Correction(mag[i], acc[i], z, P, K)
        Xs.append(X)
        ys.append(-y if flip else y)
        flips.append(flip)
    return np.array(Xs), P, np.array(ys), np.array(flips)


def trajectories():
    out = {}
    for tag, sigma, seed in (("clean", 0.0, 100), ("noisy", 0.01, 101)):
        N, T = 8, 400
        imu = make_imu(N, T, seed=seed, sigma=sigma)
        S = imu.streams.numpy()
        a0, m0 = imu.acc_ref.numpy(), imu.mag_ref.numpy()
        t_ns = np.arange(T + 1, dtype=np.int64) * 10 ** 7
        qr = [(1.0, 0.1)] * 6 + [(10.0, 0.01), (0.01, 10.0)]      # last two: other tunings
        X = np.empty((T, N, 4)); Pf = np.empty((N, 4, 4)); Y = np.empty((T, N, 4)); F = np.empty((T, N), dtype=bool)
        for n in range(N):
            X[:, n], Pf[n], Y[:, n], F[:, n] = drive(
                t_ns, S[:, 0:3, n].astype(np.float64), S[:, 3:6, n].astype(np.float64),
                S[:, 6:9, n].astype(np.float64), a0[:, n].astype(np.float64), m0[:, n].astype(np.float64), *qr[n])
        out.update({f"{tag}_streams": S, f"{tag}_acc_ref": a0, f"{tag}_mag_ref": m0, f"{tag}_X": X, f"{tag}_P": Pf,
                    f"{tag}_y": Y, f"{tag}_flips": F, f"{tag}_q": np.array([q for q, _ in qr]),
                    f"{tag}_r": np.array([r for _, r in qr])})
    out["dt"] = np.float64(0.01)
    np.savez_compressed(os.path.join(HERE, "ekf_trajectories.npz"), **out)


def edge_cases():
    """What the reference accepts beyond main_file.py's own use: a caller-supplied initial state that is not
    normalised with a full covariance, and sensors that are not normalised (accelerometer in units of 1.6 g, so that
    the weight 1 - |a_z| of ExtendedKalmanFilter.py:71 changes sign along the trajectory; magnetometer x 47)."""
    N, T = 8, 150
    imu = make_imu(N, T, seed=202, sigma=0.01)
    S = imu.streams.numpy().copy()
    a0, m0 = imu.acc_ref.numpy(), imu.mag_ref.numpy()
    t_ns = np.arange(T + 1, dtype=np.int64) * 10 ** 7
    rng = np.random.default_rng(9)
    x0 = (rng.normal(size=(N, 4)) * rng.uniform(0.5, 2.0, (N, 1))).astype(np.float32)
    M = rng.normal(size=(N, 4, 4)) * 0.3
    P0 = (M @ M.transpose(0, 2, 1) + 0.5 * np.eye(4)).astype(np.float32)
    P0 = ((P0 + P0.transpose(0, 2, 1)) / 2).astype(np.float32)
    out = {"acc_ref": a0, "mag_ref": m0, "x0": x0, "P0": P0, "dt": np.float64(0.01), "q": np.float64(1.0),
           "r": np.float64(np.float32(0.1))}
    for tag, scale_a, scale_m, use_x0 in (("state", 1.0, 1.0, True), ("sensors", 1.6, 47.0, False), ("both", 1.6, 47.0, True)):
        Sx = S.copy()
        Sx[:, 3:6] *= np.float32(scale_a)
        Sx[:, 6:9] *= np.float32(scale_m)
        X = np.empty((T, N, 4)); Pf = np.empty((N, 4, 4)); F = np.empty((T, N), dtype=bool)
        for n in range(N):
            X[:, n], Pf[n], _, F[:, n] = drive(
                t_ns, Sx[:, 0:3, n].astype(np.float64), Sx[:, 3:6, n].astype(np.float64), Sx[:, 6:9, n].astype(np.float64),
                a0[:, n].astype(np.float64), m0[:, n].astype(np.float64), float(out["q"]), float(out["r"]),
                X0=x0[n].astype(np.float64) if use_x0 else None, P0=P0[n].astype(np.float64) if use_x0 else None)
        out.update({f"{tag}_streams": Sx, f"{tag}_X": X, f"{tag}_P": Pf, f"{tag}_flips": F})
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), **out)


def wahba_cases():
    rng = np.random.default_rng(7)
    M = 1500

    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)
    acc_ref = unit(rng.normal(size=(M, 3))).astype(np.float32)
    mag_ref = unit(acc_ref * rng.uniform(-0.9, 0.9, (M, 1)) + unit(rng.normal(size=(M, 3)))).astype(np.float32)
    # measurements = references seen through a random rotation (+ small noise)
    ax = unit(rng.normal(size=(M, 3))); ang = rng.uniform(0.05, np.pi - 0.05, M)
    K = np.zeros((M, 3, 3)); K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
    Rt = np.eye(3) + np.sin(ang)[:, None, None] * K + (1 - np.cos(ang))[:, None, None] * (K @ K)
    acc = unit(np.einsum("nji,nj->ni", Rt, acc_ref) + 0.01 * rng.normal(size=(M, 3))).astype(np.float32)
    mag = unit(np.einsum("nji,nj->ni", Rt, mag_ref) + 0.01 * rng.normal(size=(M, 3))).astype(np.float32)
    res = {}
    for tag in ("half", "refw"):
        R = np.empty((M, 3, 3)); Q = np.empty((M, 4))
        ka = np.full(M, 0.5) if tag == "half" else np.abs(acc[:, 2]).astype(np.float64)
        km = np.full(M, 0.5) if tag == "half" else 1 - ka
        for n in range(M):
            w = Wahba(acc_ref[n].astype(np.float64), mag_ref[n].astype(np.float64))
            R[n] = w.getRotation(acc[n].astype(np.float64), mag[n].astype(np.float64), ka[n], km[n])
            Q[n] = w.getQuarternion(acc[n].astype(np.float64), mag[n].astype(np.float64), ka[n], km[n])
        res[f"{tag}_R"], res[f"{tag}_q"], res[f"{tag}_ka"], res[f"{tag}_km"] = R, Q, ka, km
    # RotationMatrix2Quart on its own, including the exact identity (NaN) and the reference's
    # own hand-check instance (WahbaProblem_singularValue.py:4-26 -> R = diag(-1,-1,1))
    r2q_in = np.concatenate([res["half_R"][:64], np.eye(3)[None], np.diag([-1.0, -1.0, 1.0])[None]])
    with np.errstate(all="ignore"):
        r2q_out = np.array([Wahba.RotationMatrix2Quart(m) for m in r2q_in])
    hand = Wahba(np.asarray([0.0, 0.0, 1.0]), np.asarray([-1.0, 0.0, 0.0])).getRotation(
        np.asarray([0.0, 0.0, 1.0]), np.asarray([1.0, 0.0, 0.0]), 0.5, 0.5)
    np.savez_compressed(os.path.join(HERE, "wahba_cases.npz"), acc_ref=acc_ref, mag_ref=mag_ref, acc=acc, mag=mag,
                        r2q_in=r2q_in, r2q_out=r2q_out, hand_check_R=hand, **res)


def stepwise():
    rng = np.random.default_rng(11)
    M = 64
    gyro = rng.uniform(-2, 2, (M, 3)).astype(np.float32)
    x = rng.normal(size=(M, 4)); x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    A = rng.normal(size=(M, 4, 4)) * 0.3
    P = (A @ A.transpose(0, 2, 1) + 0.05 * np.eye(4)).astype(np.float32)
    P = ((P + P.transpose(0, 2, 1)) / 2).astype(np.float32)
    dt_ns = rng.integers(5_000_000, 20_000_000, M)
    acc = rng.normal(size=(M, 3)); acc = (acc / np.linalg.norm(acc, axis=1, keepdims=True)).astype(np.float32)
    mag = rng.normal(size=(M, 3)); mag = (mag / np.linalg.norm(mag, axis=1, keepdims=True)).astype(np.float32)
    acc0 = rng.normal(size=(M, 3)); acc0 = (acc0 / np.linalg.norm(acc0, axis=1, keepdims=True)).astype(np.float32)
    mag0 = rng.normal(size=(M, 3)); mag0 = (mag0 / np.linalg.norm(mag0, axis=1, keepdims=True)).astype(np.float32)
    qs, rs = 2.0 * 1.5, 0.1        # setQ(2); setQ(1.5) -> cumulative 3.0
    z = np.empty((M, 4)); Pp = np.empty((M, 4, 4)); K = np.empty((M, 4, 4)); X = np.empty((M, 4)); Pc = np.empty((M, 4, 4))
    JA = np.empty((M, 4, 4)); JB = np.empty((M, 4, 3)); cmp_ = np.empty((M, 4)); rk = np.empty((M, 4))
    rpy = np.empty((M, 3)); nrm = np.empty(M)
    for n in range(M):
        k = KalmanFilter(1000, mag0[n].astype(np.float64), acc0[n].astype(np.float64), 0.5)
        k.setQ(2.0); k.setQ(1.5); k.setR(rs)
        g64, x64, P64 = gyro[n].astype(np.float64), x[n].astype(np.float64), P[n].astype(np.float64)
        z[n], Pp[n], K[n] = k.Prediction(g64, 1000 + int(dt_ns[n]), x64, P64)
        X[n], Pc[n] = k.Correction(mag[n].astype(np.float64), acc[n].astype(np.float64), z[n], Pp[n], K[n])
        JA[n], JB[n] = k.GetJacobian_A(g64), k.GetJacobian_B(x64)
        cmp_[n] = k.Comparator(x64, z[n])
        rk[n] = KalmanFilter.RungeKutta4(x64, int(dt_ns[n]), g64)
        rpy[n] = Quart2RPY(x64)
        nrm[n] = norm(P64[0])
    split_in = [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]
    np.savez_compressed(os.path.join(HERE, "stepwise.npz"), gyro=gyro, x=x, P=P, dt_ns=dt_ns, acc=acc, mag=mag,
                        acc0=acc0, mag0=mag0, q_scale=qs, r_scale=rs, z=z, P_pred=Pp, K=K, X=X, P_corr=Pc, JA=JA,
                        JB=JB, comparator=cmp_, rk4=rk, rpy=rpy, norm=nrm, split_in=np.array(split_in),
                        split_out=np.array(DimensionalSplit(split_in)))


def rk4_known_answer():
    """Quarternions.py:19-41 (`equation`) driven as at :99-112: omega = [3pi/2, pi, pi/2] for 1 s."""
    src = open(os.path.join(REF, "Quarternions.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def equation"))
    end = next(i for i, l in enumerate(src) if l.startswith("def RungeKutta4"))
    ns = {"np": np}
    exec("\n".join(src[start:end]), ns)
    w = np.asarray([3 * np.pi / 2, np.pi, np.pi / 2])
    res = {}
    for j in range(4):
        n_it = 10 ** j
        q = np.asarray([1.0, 0.0, 0.0, 0.0])
        for _ in range(n_it):
            q = ns["equation"](q, 1.0 / n_it, w)
        res[f"steps_{n_it}"] = q
    np.savez_compressed(os.path.join(HERE, "rk4_known_answer.npz"), omega=w, **res)
    return res


def log_format():
    """A log in the server's text format (written by poseestimationkf_b200.logio from a short reference
    run) parsed by THE REFERENCE'S OWN READER: ReadFile.getData opens a hard-coded Windows path
    (ReadFile.py:24), so `open` is redirected for that one call."""
    import builtins
    from poseestimationkf_b200 import logio
    T = 40
    imu = make_imu(1, T, seed=55, sigma=0.01)
    S = imu.streams.numpy().astype(np.float64)[:, :, 0]
    a0, m0 = imu.acc_ref.numpy()[:, 0].astype(np.float64), imu.mag_ref.numpy()[:, 0].astype(np.float64)
    t0 = 123456789000
    t_ns = t0 + (np.arange(T) + 1) * 10_000_000
    X, _, Y, _ = drive(np.concatenate([[t0], t_ns]), S[:, 0:3], S[:, 3:6], S[:, 6:9], a0, m0, 1.0, 0.1)
    qg = []
    q = np.asarray([1.0, 0.0, 0.0, 0.0])
    for i in range(T):
        q = KalmanFilter.RungeKutta4(q, 10_000_000, S[i, 0:3])
        qg.append(q)
    path = os.path.join(HERE, "sample_log.txt")
    logio.write_log(path, acc_0=a0, mag_0=m0, t0_ns=t0, t_ns=t_ns, gyro=S[:, 0:3], mag_1=S[:, 6:9], acc_1=S[:, 3:6],
                    x_k=X, wahba_quart=Y, q_gyro=qg)
    import ReadFile                                   # the reference's reader
    real_open = builtins.open

    def redirected(name, *a, **k):
        return real_open(path if str(name).endswith("KalmanFilter.txt") else name, *a, **k)
    builtins.open = redirected
    try:
        g = ReadFile.getData()
    finally:
        builtins.open = real_open
    np.savez_compressed(os.path.join(HERE, "log_parsed.npz"), mag_0=g.mag_0, mag_1=g.mag_1, acc_0=g.acc_0, acc_1=g.acc_1,
                        gyro=g.gyro, timestamp=g.timestamp, quart_wahba=g.quart_wahba, quart_xk=g.quart_xk,
                        quart_gyro=g.quart_gyro, X_full_precision=X)


if __name__ == "__main__":
    if sys.argv[1:] == ["edge"]:          # only the fixture added later; the others stay byte-identical
        edge_cases()
        sys.exit(0)
    edge_cases()
    log_format()
    trajectories()
    wahba_cases()
    stepwise()
    print(rk4_known_answer())
    print("golden fixtures written to", HERE)
