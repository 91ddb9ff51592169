"""Replay a recording in the reference's log format on the GPU and print what main_file.py plots.

    python examples/replay_log.py [path/to/KalmanFilter.txt]        (default: tests/golden/sample_log.txt)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseestimationkf_b200 import batched as B, logio   # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "sample_log.txt")
log = logio.read_log(path)
streams, acc_ref, mag_ref, dt = log.to_streams("cuda")
state, traj, flips = B.replay(streams, acc_ref, mag_ref, dt=dt, q=1.0, r=0.1, store_trajectory=True, store_flips=True)
gyro_track, wahba_track, _ = B.tracks(streams, acc_ref, mag_ref, dt=dt)
rpy = B.traj2rpy(traj)
X = traj[:, 0].cpu().numpy()
print(f"{log.n_steps()} samples from {path}")
print("filter  X_k[-1] =", np.round(X[-1], 6), " roll/pitch/yaw [deg] =", np.round(rpy[-1, 0].cpu().numpy(), 3))
print("gyro    q[-1]   =", np.round(gyro_track[-1, 0].cpu().numpy(), 6))
print("wahba   q[-1]   =", np.round(wahba_track[-1, 0].cpu().numpy(), 6))
if log.quart_xk:
    logged = np.asarray(log.quart_xk[1:])
    print("max |X_k - logged X_k| =", float(np.abs(X - logged[: len(X)]).max()), "(the log carries 6 decimals)")
print("q/-q flips:", int(flips.sum()))
