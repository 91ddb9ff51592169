"""Parity report of the GPU replay against the float64 oracle, written as one JSON document (supplementary
evidence for profiles/; the pass/fail gates are the `-m gpu` tests).  Lives under tests/ because it uses oracle/.
    python tests/parity_report.py [out.json]

  golden      the 16 trajectories frozen from the unmodified reference classes (tests/golden/ekf_trajectories.npz)
  synthetic   8192 filters x 1000 steps of the bench's generator against the compiled C oracle, every step compared
  sweep       the four corners of the 1e-3..1e3 (Q,R) grid, 256 filters x 2000 steps, precise variant
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as CO
from oracle import ekf_oracle as O
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200.synth import make_imu

dev = torch.device("cuda:0")
out = {"tolerance_rad": 1e-5, "device": torch.cuda.get_device_name(0)}


def stats(got, ref, flips=None, ref_flips=None):
    ang = O.quat_angle(got, ref)
    d = {"max_angle_rad": float(ang.max()), "mean_angle_rad": float(ang.mean()), "p9999_angle_rad": float(np.quantile(ang, 0.9999)),
         "sign_mismatches": int((np.sum(got * ref, axis=-1) <= 0).sum()), "filter_steps": int(ang.size)}
    if flips is not None:
        d["flip_mask_mismatches"] = int((flips.astype(bool) != ref_flips.astype(bool)).sum())
    return d


g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ekf_trajectories.npz"))
out["golden"] = {}
for tag in ("clean", "noisy"):
    for staging in ("ldg", "tma", "tma_packed"):
        t = lambda k: torch.from_numpy(np.ascontiguousarray(g[f"{tag}_{k}"])).to(dev)
        st, traj, flips = B.replay(t("streams"), t("acc_ref"), t("mag_ref"), dt=float(g["dt"]), q=t("q").float(), r=t("r").float(),
                                   store_trajectory=True, store_flips=True, staging=staging, precise_state=False, allow_imprecise=True)
        out["golden"][f"{tag}/{staging}"] = stats(traj.cpu().numpy(), g[f"{tag}_X"], flips.cpu().numpy(), g[f"{tag}_flips"])

N, T = 8192, 1000
imu = make_imu(N, T, seed=21, sigma=0.01, device=dev)
S = imu.streams.cpu().numpy()
ref = CO.replay(S, imu.dt * 1e9, imu.acc_ref.cpu().numpy(), imu.mag_ref.cpu().numpy(), 1.0, float(np.float32(0.1)), store=True, flips=True)
st, traj, flips = B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, q=1.0, r=0.1, store_trajectory=True, store_flips=True)
out["synthetic_8192x1000_q1_r0.1"] = stats(traj.cpu().numpy(), ref["X"], flips.cpu().numpy(), ref["flips"])
az = np.abs(S[:, 5])
out["synthetic_8192x1000_q1_r0.1"]["share_of_steps_with_|a_z|_outside_0.02_0.98"] = float(((az < 0.02) | (az > 0.98)).mean())

Ns, Ts = 256, 2000
imu = make_imu(Ns, Ts, seed=33, sigma=0.01, device=dev)
S = imu.streams.cpu().numpy()
out["sweep_corners_256x2000"] = {}
for q, r in ((1e-3, 1e3), (1e3, 1e-3), (1e-3, 1e-3), (1e3, 1e3), (1.0, 0.1)):
    ref = CO.replay(S, imu.dt * 1e9, imu.acc_ref.cpu().numpy(), imu.mag_ref.cpu().numpy(), float(np.float32(q)), float(np.float32(r)),
                    store=True, flips=False)
    for precise in (True, False):
        st, traj, _ = B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, q=q, r=r, store_trajectory=True, precise_state=precise, allow_imprecise=True)
        out["sweep_corners_256x2000"][f"q={q:g},r={r:g},{'precise' if precise else 'plain'}"] = stats(traj.cpu().numpy(), ref["X"])

text = json.dumps(out, indent=1)
print(text)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as fh:
        fh.write(text + "\n")
