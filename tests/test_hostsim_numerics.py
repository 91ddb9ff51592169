"""float32 arithmetic of the device math header (poseestimationkf_b200/csrc/ekf_math.cuh) compiled
for the CPU by tests/hostsim (a TEST-ONLY build) against the golden reference outputs.  This is the
GPU-less early warning for numerics regressions; the authoritative parity tests are the `gpu` ones."""
import numpy as np
import pytest

from oracle import ekf_oracle as O
from tests.hostsim import api as H

TOL = 1e-5   # rad, the bar stated in BASELINE.json north_star


@pytest.mark.parametrize("tag", ["clean", "noisy"])
def test_replay_f32_qr2_vs_reference(golden_traj, tag):
    g = golden_traj
    traj, flips, P = H.replay(g[f"{tag}_streams"], 0.01, g[f"{tag}_acc_ref"], g[f"{tag}_mag_ref"], g[f"{tag}_q"],
                              g[f"{tag}_r"], precision="f32", algo="qr2")
    got = traj.transpose(0, 2, 1)
    ang = O.quat_angle(got, g[f"{tag}_X"])
    assert ang.max() < TOL, ang.max()
    assert (np.sum(got * g[f"{tag}_X"], axis=-1) > 0).all()       # same q/-q branch everywhere
    assert (flips == g[f"{tag}_flips"]).all()
    # final covariance (upper triangle) against the reference's P
    tri = [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)]
    ref = np.stack([g[f"{tag}_P"][:, i, j] for i, j in tri])
    np.testing.assert_allclose(P, ref, rtol=2e-4, atol=2e-6)


def test_replay_f64_restructured_algebra_is_exact(golden_traj):
    # the algebraic restructurings (RK4 polynomial, K = I - r S^-1, P = r K, rank-2 Wahba) are exact:
    # in float64 they reproduce the reference to ~1e-10 (q and r reach the kernel as float32: 0.1f != 0.1)
    g = golden_traj
    for algo in ("qr2", "jacobi"):
        traj, flips, _ = H.replay(g["noisy_streams"], 0.01, g["noisy_acc_ref"], g["noisy_mag_ref"], g["noisy_q"],
                                  g["noisy_r"], precision="f64", algo=algo)
        assert O.quat_angle(traj.transpose(0, 2, 1), g["noisy_X"]).max() < 1e-9
        assert (flips == g["noisy_flips"]).all()


def _wahba_condition(g, tag):
    """sigma1/sigma2 of the reference's B matrix per golden case (its float32 formation costs
    ~eps32*sigma1/sigma2 in the rotation -- SURVEY.md section 7, hard part 1)."""
    B = (g[f"{tag}_ka"][:, None, None] * g["acc_ref"][:, :, None].astype(np.float64) * g["acc"][:, None, :]
         + g[f"{tag}_km"][:, None, None] * g["mag_ref"][:, :, None].astype(np.float64) * g["mag"][:, None, :])
    s = np.linalg.svd(B, compute_uv=False)
    return s[:, 0] / s[:, 1]


@pytest.mark.parametrize("tag,algo", [("half", "qr2"), ("half", "jacobi"), ("refw", "qr2"), ("refw", "jacobi")])
def test_wahba_f32(golden_wahba, tag, algo):
    g = golden_wahba
    q, R = H.wahba(g["acc_ref"].T, g["mag_ref"].T, g["acc"].T, g["mag"].T, g[f"{tag}_ka"], g[f"{tag}_km"],
                   precision="f32", algo=algo, sweeps=5, want_R=True)
    ang = O.quat_angle(q.T, g[f"{tag}_q"])
    cond = _wahba_condition(g, tag)
    if algo == "qr2":
        # the rank-2 form never builds B: accurate regardless of sigma1/sigma2 (up to 2.6e4 in this set)
        assert ang.max() < 2e-6, ang.max()
        np.testing.assert_allclose(R.T.reshape(-1, 3, 3), g[f"{tag}_R"], atol=5e-6)
    else:
        # B formed in float32 as the reference forms it: error bounded by the conditioning of B
        assert (ang < 1e-6 + 4 * 6e-8 * cond).all(), (ang / (1e-6 + 4 * 6e-8 * cond)).max()
        assert np.median(ang) < 3e-7
    # the sign convention of the 3-branch conversion (component of the winning branch positive) is
    # reproduced except at near-ties of the branch traces
    same = np.sum(q.T * g[f"{tag}_q"], axis=1) > 0
    assert (~same).sum() <= 2


def test_lowpass_and_lpf_replay_match_oracle(golden_traj):
    g = golden_traj
    S = g["noisy_streams"]
    T, _, N = S.shape
    traj, _, _ = H.replay(S, 0.01, g["noisy_acc_ref"], g["noisy_mag_ref"], 1.0, 0.1, precision="f32", algo="qr2",
                          lpf_acc=0.1, lpf_mag=0.1)
    # oracle: low-pass in float64 first (state from 0), then the reference filter on the filtered values
    Sf = S.astype(np.float64).copy()
    for n in range(N):
        Sf[:, 3:6, n] = O.lowpass_scalar(S[:, 3:6, n], 0.1)
        Sf[:, 6:9, n] = O.lowpass_scalar(S[:, 6:9, n], 0.1)
    ref = O.replay_batched(np.full(T, 1e7), Sf[:, 0:3], Sf[:, 3:6], Sf[:, 6:9], g["noisy_acc_ref"].T.astype(np.float64),
                           g["noisy_mag_ref"].T.astype(np.float64), 1.0, 0.1)
    # the first steps start from a zero low-pass state: tiny vectors but well-defined directions
    assert O.quat_angle(traj.transpose(0, 2, 1), ref["X"]).max() < TOL
