"""Freezes golden vectors for the raw-sensor pre-processing rows from the reference's OWN C++,
compiled by oracle/Makefile into oracle/_ref/libposekf_ref.so (InitialValues.cpp as it lies in the reference tree and two
functions of Parser.cpp extracted at build time).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_cpp.py [out_dir]

Writes preprocess_ref.npz (interpolate + normalise, Parser.cpp:221-228,259-267 as sequenced by :232-242) and
initial_values_ref.npz (InitialValues.cpp:19-66 + the normalisation of Parser.cpp:46-49).  The GPU box never needs the
reference: the tests read these fixtures."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
out_dir = sys.argv[1] if len(sys.argv) > 1 else HERE
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libposekf_ref.so"))
dp = C.POINTER(C.c_double)
lib.ref_interpolate_normalise.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, dp, dp, dp]
lib.ref_interpolate.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, dp, dp, dp]
lib.ref_normalize.argtypes = [dp]
lib.ref_initial_values.argtypes = [dp, C.c_int, dp, dp]
lib.ref_initial_values.restype = C.c_int


def p(a):
    return a.ctypes.data_as(dp)


rng = np.random.default_rng(20261019)

# ---- interpolate + normalise: M cases; sensor values are float32-representable (what the kernels are fed) -------------
M = 4096
scale = np.concatenate([np.full(M // 2, 9.81), np.full(M - M // 2, 45.0)])        # accelerometer m/s^2, magnetometer uT
y1 = (rng.normal(size=(M, 3)) * scale[:, None]).astype(np.float32).astype(np.float64)
y2 = (y1 + rng.normal(size=(M, 3)) * 0.05 * scale[:, None]).astype(np.float32).astype(np.float64)
t1 = rng.integers(10 ** 12, 10 ** 13, M)                                          # absolute nanosecond stamps
span = rng.integers(2 * 10 ** 6, 2 * 10 ** 7, M)                                  # 2..20 ms between the two samples
t2 = t1 + span
t3 = t1 + (span * rng.uniform(0.0, 1.0, M)).astype(np.int64)                      # gyro stamp between them ...
t3[:64] = t1[:64]                                                                 # ... or exactly on the first / second sample,
t3[64:128] = t2[64:128]
t3[128:192] = t2[128:192] + span[128:192] // 3                                    # ... or past the second (extrapolation)
interp = np.empty((M, 3))
unit = np.empty((M, 3))
for i in range(M):
    a, b = np.ascontiguousarray(y1[i]), np.ascontiguousarray(y2[i])
    o = np.empty(3)
    lib.ref_interpolate(int(t1[i]), int(t2[i]), int(t3[i]), p(a), p(b), p(o))
    interp[i] = o
    lib.ref_interpolate_normalise(int(t1[i]), int(t2[i]), int(t3[i]), p(a), p(b), p(o))
    unit[i] = o
np.savez_compressed(os.path.join(out_dir, "preprocess_ref.npz"), y1=y1, y2=y2, t1=t1, t2=t2, t3=t3, interpolated=interp,
                    normalised=unit)

# ---- initial values: N sensors x K samples ------------------------------------------------------------------------------
N, K = 256, 100                                                                   # Parser.cpp:5-7: InitialValues(100)
mean_true = rng.normal(size=(N, 3)) * np.concatenate([np.full(N // 2, 9.81), np.full(N - N // 2, 45.0)])[:, None]
samples = (mean_true[:, None, :] + rng.normal(size=(N, K, 3)) * 0.05 * np.abs(mean_true).max(axis=1)[:, None, None])
samples = samples.astype(np.float32).astype(np.float64)
avg, var, avg_unit = np.empty((N, 3)), np.empty((N, 3)), np.empty((N, 3))
for n in range(N):
    s = np.ascontiguousarray(samples[n])
    a, v = np.empty(3), np.empty(3)
    assert lib.ref_initial_values(p(s), K, p(a), p(v)) == 0
    avg[n], var[n] = a, v
    u = a.copy()
    lib.ref_normalize(p(u))                                                        # Parser.cpp:48-49
    avg_unit[n] = u
np.savez_compressed(os.path.join(out_dir, "initial_values_ref.npz"), samples=samples, avg=avg, var=var, avg_unit=avg_unit)
print("wrote preprocess_ref.npz, initial_values_ref.npz to", out_dir)
