"""The compiled C oracle (oracle/ekf_oracle.c) against the frozen outputs of the unmodified reference
and against the pinned numpy oracle.  CPU only."""
import time

import numpy as np

from oracle import c_oracle as CO
from oracle import ekf_oracle as O
from poseestimationkf_b200.synth import make_imu


def test_c_oracle_vs_reference_golden(golden_traj):
    g = golden_traj
    for tag in ("clean", "noisy"):
        out = CO.replay(g[f"{tag}_streams"], 1e7, g[f"{tag}_acc_ref"], g[f"{tag}_mag_ref"], g[f"{tag}_q"], g[f"{tag}_r"])
        assert O.quat_angle(out["X"], g[f"{tag}_X"]).max() < 1e-11     # near rank-1 Wahba steps: eps64*sigma1/sigma2
        assert (np.sum(out["X"] * g[f"{tag}_X"], axis=-1) > 0).all()
        assert (out["flips"] == g[f"{tag}_flips"]).all()
        np.testing.assert_allclose(out["P_final"], g[f"{tag}_P"], rtol=1e-9, atol=1e-15)


def test_c_oracle_wahba_vs_reference_golden(golden_wahba):
    g = golden_wahba
    for tag in ("half", "refw"):
        R, q = CO.wahba(g["acc_ref"].T, g["mag_ref"].T, g["acc"].T, g["mag"].T, g[f"{tag}_ka"], g[f"{tag}_km"])
        np.testing.assert_allclose(R, g[f"{tag}_R"], atol=2e-10)     # sigma1/sigma2 up to 2.6e4 in this set
        assert O.quat_angle(q, g[f"{tag}_q"]).max() < 2e-10
        assert (np.sum(q * g[f"{tag}_q"], axis=1) > 0).all()


def test_c_oracle_vs_numpy_oracle_and_speed():
    imu = make_imu(256, 400, seed=17, sigma=0.01)
    S = imu.streams.numpy()
    ref = O.replay_batched(np.full(400, 1e7), S[:, 0:3], S[:, 3:6], S[:, 6:9], imu.acc_ref.numpy().T,
                           imu.mag_ref.numpy().T, 1.0, 0.1)
    t0 = time.perf_counter()
    out = CO.replay(S, 1e7, imu.acc_ref.numpy(), imu.mag_ref.numpy(), 1.0, 0.1)
    dt = time.perf_counter() - t0
    assert O.quat_angle(out["X"], ref["X"]).max() < 1e-11
    assert (out["flips"] == ref["flips"]).all()
    assert 256 * 400 / dt > 1e4          # orders of magnitude faster than the numpy form: usable as a bulk checker
