"""Per-step latency of the reference-named drop-in modules at batch=1 (the loop of main_file.py:38-47)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseestimationkf_b200 import compat
from poseestimationkf_b200.synth import make_imu
sys.path.insert(0, compat.PATH)
from ExtendedKalmanFilter import KalmanFilter
from Wahba import Wahba
T = 300
imu = make_imu(1, T, seed=2, sigma=0.01)
S = imu.streams.numpy().astype(np.float64)[:, :, 0]
a0, m0 = imu.acc_ref.numpy()[:, 0].astype(np.float64), imu.mag_ref.numpy()[:, 0].astype(np.float64)
t_ns = np.arange(T + 1, dtype=np.int64) * 10 ** 7
k = KalmanFilter(t_ns[0], m0, a0, 0.5); k.setQ(1); k.setR(0.1)
w = Wahba(a0, m0)
P = np.identity(4); X = np.asarray([1., 0., 0., 0.])
for rep in range(2):
    t0 = time.perf_counter()
    for i in range(T):
        z, P, K = k.Prediction(S[i, 0:3], t_ns[i + 1] + rep * 10 ** 10, X, P)
        X, P = k.Correction(S[i, 6:9], S[i, 3:6], z, P, K)
    dt = time.perf_counter() - t0
print("compat loop: %.1f us per filter-step (Prediction+Correction), batch=1" % (dt / T * 1e6))
# the same loop through the reference's own numpy classes where its tree is present (POSEKF_REF or /root/reference),
# for the comparison VERDICT r01 item 9 asks for; on the GPU box compare with bench.py's cpu_baseline.single_core_value
ref_dir = os.environ.get("POSEKF_REF", "/root/reference/Python Kalman Filter")
if os.path.exists(os.path.join(ref_dir, "ExtendedKalmanFilter.py")):
    for m in ("ExtendedKalmanFilter", "Wahba", "UtilityFunctions"):
        sys.modules.pop(m, None)
    sys.path.insert(0, ref_dir)
    from ExtendedKalmanFilter import KalmanFilter as RefKF
    rk = RefKF(t_ns[0], m0, a0, 0.5); rk.setQ(1); rk.setR(0.1)
    P = np.identity(4); X = np.asarray([1., 0., 0., 0.])
    t0 = time.perf_counter()
    for i in range(T):
        z, P, K = rk.Prediction(S[i, 0:3], t_ns[i + 1], X, P)
        X, P = rk.Correction(S[i, 6:9], S[i, 3:6], z, P, K)
    print("reference numpy loop: %.1f us per filter-step" % ((time.perf_counter() - t0) / T * 1e6))
t0 = time.perf_counter()
for i in range(T):
    w.getQuarternion(S[i, 3:6], S[i, 6:9], 0.5, 0.5)
print("Wahba.getQuarternion: %.1f us per call" % ((time.perf_counter() - t0) / T * 1e6))
