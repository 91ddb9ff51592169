"""Developer probe (CPU, no GPU): float32 accuracy of the device math header under experiment flags.

    python tools/hostsim_accuracy.py [-DPKF_FUSE=1 ...]

Builds tests/hostsim/hostsim.cpp with the given -D flags into a scratch library and reports the worst quaternion
angle against (a) the frozen reference trajectories and (b) the compiled float64 oracle on a synthetic batch at a
few (Q,R) tunings inside the plain variant's range.  Test tooling only (uses oracle/ and tests/hostsim)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as CO
from oracle import ekf_oracle as O
from poseestimationkf_b200.synth import make_imu
from tests.hostsim import api as H

flags = [a for a in sys.argv[1:] if a.startswith("-D")]
so = os.path.join(tempfile.mkdtemp(), "libhostsim_x.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-ffp-contract=off", "-fPIC", "-shared", *flags, "-o", so,
                       os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
H._lib = C.CDLL(so)

g = np.load(os.path.join(ROOT, "tests", "golden", "ekf_trajectories.npz"))
out = {"flags": flags}
for tag in ("clean", "noisy"):
    traj, flips, _ = H.replay(g[f"{tag}_streams"], 0.01, g[f"{tag}_acc_ref"], g[f"{tag}_mag_ref"], g[f"{tag}_q"], g[f"{tag}_r"])
    ang = O.quat_angle(traj.transpose(0, 2, 1), g[f"{tag}_X"])
    out[f"golden_{tag}"] = {"max": float(ang.max()), "mean": float(ang.mean()), "flip_mismatch": int((flips != g[f"{tag}_flips"]).sum())}

N, T = 512, 1000
imu = make_imu(N, T, seed=4242, sigma=0.01)
S = imu.streams.numpy()
ar, mr = imu.acc_ref.numpy(), imu.mag_ref.numpy()
for q, r in ((1.0, 0.1), (1.0, 50.0), (1e3, 0.2), (0.01, 0.9), (100.0, 100.0)):
    ref = CO.replay(S, imu.dt * 1e9, ar, mr, float(np.float32(q)), float(np.float32(r)))
    for name, fn in (("scalar", lambda: H.replay(S, imu.dt, ar, mr, q, r)[:2]),
                     ("packed", lambda: (lambda t, p, f: (t, f))(*H.replay_packed(S, imu.dt, ar, mr, q, r)))):
        traj, flips = fn()
        ang = O.quat_angle(traj.transpose(0, 2, 1), ref["X"])
        out[f"synth_q{q}_r{r}_{name}"] = {"max": float(ang.max()), "mean": float(ang.mean()),
                                           "flip_mismatch": int((flips != ref["flips"]).sum())}
for k, v in out.items():
    print(k, v)
