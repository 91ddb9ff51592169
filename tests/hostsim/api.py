"""ctypes front end of the test-only CPU build of the device math header (see hostsim.cpp)."""
import ctypes as C

import numpy as np

from .build import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def replay(streams, dt, acc_ref, mag_ref, q, r, *, precision="f32", algo="qr2", lpf_acc=-1.0, lpf_mag=-1.0,
           compensated=False, x0=None, p0_over_r=None):
    """streams [T,9,N] f32, acc_ref/mag_ref [3,N] f32, q/r [N] f32, dt scalar or [T] f32 (seconds).
    x0 [4,N] / p0_over_r [10,N] (upper triangle of P0/r): initial state as the kernels' state buffers hold it.
    Returns traj [T,4,N] f64, flips [T,N] bool, P [10,N] f64."""
    streams = np.ascontiguousarray(streams, dtype=np.float32)
    T, _, N = streams.shape
    dt = np.atleast_1d(np.asarray(dt, dtype=np.float64))   # f32 build rounds it to float32 itself
    acc_ref = np.ascontiguousarray(acc_ref, dtype=np.float32)
    mag_ref = np.ascontiguousarray(mag_ref, dtype=np.float32)
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float32), (N,)))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float32), (N,)))
    traj = np.empty((T, 4, N))
    flips = np.empty((T, N), dtype=np.uint8)
    P = np.empty((10, N))
    x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float32)
    p0 = None if p0_over_r is None else np.ascontiguousarray(p0_over_r, dtype=np.float32)
    rc = lib().hostsim_replay(C.c_int(0 if precision == "f32" else 1), C.c_int(0 if algo == "qr2" else 1),
                              C.c_int(int(compensated)), C.c_int64(N), C.c_int64(T), _p(streams, C.c_float), _p(dt, C.c_double),
                              C.c_int(int(dt.size > 1)), _p(acc_ref, C.c_float), _p(mag_ref, C.c_float),
                              _p(q, C.c_float), _p(r, C.c_float), C.c_float(lpf_acc), C.c_float(lpf_mag),
                              _p(traj, C.c_double), _p(flips, C.c_uint8), _p(P, C.c_double), _p(x0, C.c_float), _p(p0, C.c_float))
    assert rc == 0
    return traj, flips.astype(bool), P


def wahba(acc_ref, mag_ref, acc, mag, ka, km, *, precision="f32", algo="qr2", sweeps=5, want_R=False):
    """all vectors [3,N] f32, weights [N] f32 -> q [4,N] f64 (and R [9,N])."""
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (acc_ref, mag_ref, acc, mag)]
    N = arrs[0].shape[1]
    ka = np.ascontiguousarray(np.broadcast_to(np.asarray(ka, dtype=np.float32), (N,)))
    km = np.ascontiguousarray(np.broadcast_to(np.asarray(km, dtype=np.float32), (N,)))
    q = np.empty((4, N))
    R = np.empty((9, N)) if want_R else None
    rc = lib().hostsim_wahba(C.c_int(0 if precision == "f32" else 1), C.c_int({"qr2": 0, "jacobi": 1, "quat2": 2}[algo]),
                             C.c_int(sweeps), C.c_int64(N), *[_p(a, C.c_float) for a in arrs],
                             _p(ka, C.c_float), _p(km, C.c_float), _p(q, C.c_double), _p(R, C.c_double))
    assert rc == 0
    return (q, R) if want_R else q


def replay_packed(streams, dt, acc_ref, mag_ref, q, r, *, lpf_acc=-1.0, lpf_mag=-1.0, compensated=False):
    """The packed (two filters per thread, f32x2) form of the float32 replay.  Same arguments / results as
    `replay` (without the flip mask); N must be even."""
    streams = np.ascontiguousarray(streams, dtype=np.float32)
    T, _, N = streams.shape
    dt = np.atleast_1d(np.asarray(dt, dtype=np.float64))
    acc_ref = np.ascontiguousarray(acc_ref, dtype=np.float32)
    mag_ref = np.ascontiguousarray(mag_ref, dtype=np.float32)
    q = np.ascontiguousarray(np.broadcast_to(np.asarray(q, dtype=np.float32), (N,)))
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float32), (N,)))
    traj = np.empty((T, 4, N))
    P = np.empty((10, N))
    flips = np.empty((T, N), dtype=np.uint8)
    rc = lib().hostsim_replay_packed(C.c_int(int(compensated)), C.c_int64(N), C.c_int64(T), _p(streams, C.c_float),
                                     _p(dt, C.c_double), C.c_int(int(dt.size > 1)), _p(acc_ref, C.c_float),
                                     _p(mag_ref, C.c_float), _p(q, C.c_float), _p(r, C.c_float), C.c_float(lpf_acc),
                                     C.c_float(lpf_mag), _p(traj, C.c_double), _p(P, C.c_double), _p(flips, C.c_uint8))
    assert rc == 0
    return traj, P, flips.astype(bool)
