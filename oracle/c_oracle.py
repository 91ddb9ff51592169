"""ctypes front end of oracle/libekf_oracle.so (C float64 restatement) -- TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libekf_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "ekf_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = C.CDLL(SO)
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def replay(streams, dt_ns, acc_ref, mag_ref, q, r, store=True, flips=True, threads=None):
    """streams [T,9,N] float32; dt_ns [T]; acc_ref/mag_ref [3,N] float32; q, r scalars or [N].
    Returns dict(X [T,N,4] or None, X_final [N,4], P_final [N,4,4], flips [T,N] or None)."""
    streams = np.ascontiguousarray(streams, dtype=np.float32)
    T, _, N = streams.shape
    dt_ns = np.ascontiguousarray(np.broadcast_to(np.asarray(dt_ns, dtype=np.float64), (T,)))
    acc_ref = np.ascontiguousarray(acc_ref, dtype=np.float32)
    mag_ref = np.ascontiguousarray(mag_ref, dtype=np.float32)
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float64), (N,)))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (N,)))
    traj = np.empty((T, N, 4)) if store else None
    x = np.empty((N, 4))
    P = np.empty((N, 16))
    fl = np.empty((T, N), dtype=np.uint8) if flips else None
    rc = lib().oracle_replay_f64(C.c_int64(N), C.c_int64(T), _p(streams, C.c_float), _p(dt_ns, C.c_double),
                                 _p(acc_ref, C.c_float), _p(mag_ref, C.c_float), _p(q, C.c_double), _p(r, C.c_double),
                                 _p(traj, C.c_double), _p(x, C.c_double), _p(P, C.c_double), _p(fl, C.c_uint8),
                                 C.c_int(threads or os.cpu_count() or 1))
    assert rc == 0
    return dict(X=traj, X_final=x, P_final=P.reshape(N, 4, 4), flips=None if fl is None else fl.astype(bool))


def wahba(acc_ref, mag_ref, acc, mag, ka, km):
    """[3,N] float32 inputs, [N] weights -> (R [N,3,3], q [N,4])"""
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (acc_ref, mag_ref, acc, mag)]
    N = arrs[0].shape[1]
    ka = np.ascontiguousarray(np.broadcast_to(np.asarray(ka, dtype=np.float64), (N,)))
    km = np.ascontiguousarray(np.broadcast_to(np.asarray(km, dtype=np.float64), (N,)))
    R = np.empty((N, 9))
    q = np.empty((N, 4))
    rc = lib().oracle_wahba_f64(C.c_int64(N), *[_p(a, C.c_float) for a in arrs], _p(ka, C.c_double), _p(km, C.c_double),
                                _p(R, C.c_double), _p(q, C.c_double))
    assert rc == 0
    return R.reshape(N, 3, 3), q
