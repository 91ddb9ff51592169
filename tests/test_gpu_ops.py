"""Parity of the stand-alone operators and of the reference-named compat classes (GPU)."""
import sys

import numpy as np
import pytest
import torch

from oracle import ekf_oracle as O
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200 import compat

pytestmark = pytest.mark.gpu


def _dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)


def _cond(g, tag):
    Bm = (g[f"{tag}_ka"][:, None, None] * g["acc_ref"][:, :, None].astype(np.float64) * g["acc"][:, None, :]
          + g[f"{tag}_km"][:, None, None] * g["mag_ref"][:, :, None].astype(np.float64) * g["mag"][:, None, :])
    s = np.linalg.svd(Bm, compute_uv=False)
    return s[:, 0] / s[:, 1]


@pytest.mark.parametrize("tag", ["half", "refw"])
@pytest.mark.parametrize("algo", ["qr2", "jacobi"])
def test_wahba_vs_reference_golden(golden_wahba, cuda, tag, algo):
    g = golden_wahba
    args = [_dev(g[k].T, cuda) for k in ("acc_ref", "mag_ref", "acc", "mag")]
    R, q = B.wahba(*args, k_acc=_dev(g[f"{tag}_ka"], cuda), k_mag=_dev(g[f"{tag}_km"], cuda), want_rotation=True,
                   algo=algo)
    ang = O.quat_angle(q.cpu().numpy().T, g[f"{tag}_q"])
    # both solvers hold the reference's rotation for BOTH weightings (the Jacobi SVD is QR-preconditioned: it works on the
    # 2x2 core of the rank-2 problem, so the reference's near-rank-1 weights cost it nothing -- wahba_jacobi in ekf_math.cuh)
    assert ang.max() < (2e-6 if algo == "qr2" else 4e-6), ang.max()
    np.testing.assert_allclose(R.cpu().numpy().T.reshape(-1, 3, 3), g[f"{tag}_R"], atol=5e-6 if algo == "qr2" else 1e-5)
    assert (np.sum(q.cpu().numpy().T * g[f"{tag}_q"], axis=1) < 0).sum() == 0       # the reference's sign convention, every case
    if tag == "half":       # scalar weights and the reference-weights shortcut go through the same kernel
        _, q2 = B.wahba(*args, k_acc=0.5, k_mag=0.5, algo=algo)
        assert torch.equal(q2, q)
    else:
        _, q3 = B.wahba(*args, weights_from_acc=True, algo=algo)
        assert O.quat_angle(q3.cpu().numpy().T, g["refw_q"]).max() < (2e-6 if algo == "qr2" else 4e-6)


def test_wahba_hand_check_and_shared_reference(cuda):
    # the reference's own known answer (WahbaProblem_singularValue.py): R = diag(-1,-1,1)
    ra, rm = _dev([0.0, 0.0, 1.0], cuda), _dev([-1.0, 0.0, 0.0], cuda)
    acc, mag = _dev([[0.0], [0.0], [1.0]], cuda), _dev([[1.0], [0.0], [0.0]], cuda)
    for algo in ("qr2", "jacobi"):
        R, _ = B.wahba(ra, rm, acc, mag, k_acc=0.5, k_mag=0.5, want_rotation=True, algo=algo)
        np.testing.assert_allclose(R.cpu().numpy().reshape(3, 3), np.diag([-1.0, -1.0, 1.0]), atol=1e-6)


def test_rot2quat_including_identity_nan(golden_wahba, cuda):
    g = golden_wahba
    out = B.rot2quat(_dev(g["r2q_in"].reshape(-1, 9).T, cuda)).cpu().numpy().T
    ref = g["r2q_out"]
    ok = np.isfinite(ref).all(axis=1)
    assert O.quat_angle(out[ok], ref[ok]).max() < 2e-6
    assert (np.sum(out[ok] * ref[ok], axis=1) > 0).all()          # same sign convention
    ident = out[64]                                                 # M == I -> [nan, nan, nan, 0] like the reference
    assert np.isnan(ident[:3]).all() and ident[3] == 0.0
    np.testing.assert_allclose(np.abs(out[65]), [0, 0, 0, 1], atol=1e-6)   # diag(-1,-1,1): 180 deg about z
    np.testing.assert_allclose(out[ok], ref[ok], rtol=0, atol=2e-6)        # component by component, not only as a rotation
    # what the reference accepts, the operator accepts: scaled / non-orthogonal matrices give the reference's
    # (un-normalised) vector -- the oracle's restatement is pinned bit for bit to RotationMatrix2Quart (tests/test_oracle.py)
    rng = np.random.default_rng(11)
    M = (g["r2q_in"][:48] * rng.uniform(0.3, 3.0, (48, 1, 1)) + 0.05 * rng.normal(size=(48, 3, 3))).astype(np.float32)
    want = np.array([O.rotation_to_quat(m.astype(np.float64)) for m in M])
    got = B.rot2quat(_dev(M.reshape(-1, 9).T, cuda)).cpu().numpy().T
    fin = np.isfinite(want).all(axis=1)
    assert fin.sum() >= 40
    np.testing.assert_allclose(got[fin], want[fin], rtol=3e-7, atol=3e-7)
    assert (np.abs(np.linalg.norm(want[fin], axis=1) - 1) > 0.05).any()     # genuinely un-normalised cases


def test_stepwise_operators_vs_reference_golden(golden_step, cuda):
    g = golden_step
    M = g["x"].shape[0]
    gyro, x = _dev(g["gyro"].T, cuda), _dev(g["x"].T, cuda)
    P = _dev(g["P"].reshape(M, 16).T, cuda)
    dt = _dev(g["dt_ns"] * 1e-9, cuda)
    qm = _dev(np.identity(3).reshape(-1) * float(g["q_scale"]), cuda)
    rm = _dev(np.identity(4).reshape(-1) * float(g["r_scale"]), cuda)
    z, Pp, K = B.predict(gyro, dt, x, P, qm, rm)
    assert O.quat_angle(z.cpu().numpy().T, g["z"]).max() < 1e-6
    np.testing.assert_allclose(Pp.cpu().numpy().T.reshape(M, 4, 4), g["P_pred"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(K.cpu().numpy().T.reshape(M, 4, 4), g["K"], rtol=1e-3, atol=1e-4)
    # Correction fed with the reference's own z, P, K so that only this operator is under test
    X, Pc, flip, meas = B.correct(_dev(g["mag"].T, cuda), _dev(g["acc"].T, cuda), _dev(g["acc0"].T, cuda),
                                  _dev(g["mag0"].T, cuda), _dev(g["z"].T, cuda), _dev(g["P_pred"].reshape(M, 16).T, cuda),
                                  _dev(g["K"].reshape(M, 16).T, cuda), want_flip=True, want_meas=True)
    assert O.quat_angle(X.cpu().numpy().T, g["X"]).max() < 2e-6
    assert (np.sum(X.cpu().numpy().T * g["X"], axis=1) > 0).all()
    np.testing.assert_allclose(Pc.cpu().numpy().T.reshape(M, 4, 4), g["P_corr"], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(B.jacobian_a(gyro).cpu().numpy().T.reshape(M, 4, 4), g["JA"], rtol=1e-7)
    np.testing.assert_allclose(B.jacobian_b(x).cpu().numpy().T.reshape(M, 4, 3), g["JB"], rtol=1e-7)
    np.testing.assert_allclose(B.comparator(x, _dev(g["z"].T, cuda)).cpu().numpy().T, g["comparator"], atol=1e-6)
    assert O.quat_angle(B.rk4(x, dt, gyro).cpu().numpy().T, g["rk4"]).max() < 1e-6
    np.testing.assert_allclose(B.quat2rpy(x).cpu().numpy().T, g["rpy"], atol=1e-4)      # degrees (float32 output: 1.5e-5 at 180)
    np.testing.assert_allclose(B.norm(_dev(g["P"][:, 0, :].T, cuda)).cpu().numpy(), g["norm"], rtol=1e-6)


def test_rk4_known_answer(golden_rk4, cuda):
    # Quarternions.py: omega=[3pi/2,pi,pi/2] rad/s for 1 s -> converges to the analytic exponential
    w = _dev(golden_rk4["omega"][:, None], cuda)
    for n_it in (1, 10, 100, 1000):
        q = _dev([[1.0], [0.0], [0.0], [0.0]], cuda)
        for _ in range(n_it):
            q = B.rk4(q, 1.0 / n_it, w)
        assert O.quat_angle(q.cpu().numpy().T, golden_rk4[f"steps_{n_it}"][None]).max() < (3e-5 if n_it == 1000 else 2e-6)


def test_lowpass_operator(cuda):
    x = torch.randn((200, 3, 300), device=cuda)
    y, state = B.lowpass(x, 0.1)
    ref = np.stack([O.lowpass_scalar(x[:, :, n].cpu().numpy(), 0.1) for n in range(0, 300, 37)], axis=-1)
    np.testing.assert_allclose(y[:, :, ::37].cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
    assert torch.equal(state, y[-1])
    # chunked low-pass carries its state
    y1, s1 = B.lowpass(x[:77].contiguous(), 0.1)
    y2, _ = B.lowpass(x[77:].contiguous(), 0.1, state=s1)
    assert torch.equal(torch.cat([y1, y2]), y)


def test_compat_modules_run_the_reference_loop(golden_traj, cuda):
    """The loop body of Python Kalman Filter/main_file.py:19-47, verbatim, on the drop-in modules."""
    sys.path.insert(0, compat.PATH)
    try:
        from ExtendedKalmanFilter import KalmanFilter
        from Wahba import Wahba
        from UtilityFunctions import DimensionalSplit, norm, Quart2RPY
        g = golden_traj
        n, T = 2, 60
        S = g["noisy_streams"].astype(np.float64)
        t_ns = np.arange(T + 1, dtype=np.int64) * 10 ** 7
        acc_0, mag_0 = g["noisy_acc_ref"][:, n].astype(np.float64), g["noisy_mag_ref"][:, n].astype(np.float64)
        w = Wahba(acc_0, mag_0)
        k = KalmanFilter(t_ns[0], mag_0, acc_0, 0.5)
        k.setQ(1)
        k.setR(0.1)
        P = np.identity(4)
        X = np.asarray([1., 0., 0., 0.])
        X_k = [X]
        for i in range(T):
            z_k, P, K_k = k.Prediction(S[i, 0:3, n], t_ns[i + 1], X, P)
            wahbaquart = w.getQuarternion(S[i, 3:6, n], S[i, 6:9, n], 0.5, 0.5)
            X, P = k.Correction(S[i, 6:9, n], S[i, 3:6, n], z_k, P, K_k)
            X_k.append(X)
        got = np.array(X_k[1:])
        assert O.quat_angle(got, g["noisy_X"][:T, n]).max() < 1e-5
        assert wahbaquart.shape == (4,) and abs(norm(wahbaquart) - 1) < 1e-6
        Filt = DimensionalSplit(X_k)
        assert len(Filt) == 4 and len(Filt[0]) == T + 1
        assert k.Q[0, 0] == 1.0 and abs(k.R[0, 0] - 0.1) < 1e-15 and k.previousT == t_ns[T]
        k.setQ(2); k.setQ(3)
        assert k.Q[1, 1] == 6.0                                     # cumulative, like the reference
        np.testing.assert_allclose(Quart2RPY(X), O.quat_to_rpy_deg(X), atol=1e-4)
        R = w.getRotation(S[5, 3:6, n], S[5, 6:9, n], 0.5, 0.5)
        np.testing.assert_allclose(R @ R.T, np.identity(3), atol=1e-5)
        np.testing.assert_allclose(Wahba.RotationMatrix2Quart(R), w.getQuarternion(S[5, 3:6, n], S[5, 6:9, n], 0.5, 0.5), atol=1e-6)
        assert k.GetJacobian_A(S[0, 0:3, n]).shape == (4, 4) and k.GetJacobian_B(X).shape == (4, 3)
        np.testing.assert_allclose(KalmanFilter.RungeKutta4(X, 10 ** 7, S[0, 0:3, n]), O.rk4(X, 10 ** 7, S[0, 0:3, n]), atol=1e-6)
        assert abs(k.Comparator(X, X)[0] - 1.0) < 1e-6
    finally:
        sys.path.remove(compat.PATH)
        for m in ("ExtendedKalmanFilter", "Wahba", "UtilityFunctions", "_bridge"):
            sys.modules.pop(m, None)


def test_comparison_tracks_and_tuning_objective(cuda):
    """gyro-only + Wahba-only tracks, RPY of a trajectory, on-device loss of a sweep."""
    from poseestimationkf_b200.synth import make_imu
    Ns, T = 256, 200
    imu = make_imu(Ns, T, seed=31, sigma=0.01, device=cuda, keep_truth=True)
    gyro, wah, gstate = B.tracks(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt)
    S = imu.streams.cpu().numpy().astype(np.float64)
    a0, m0 = imu.acc_ref.cpu().numpy().T.astype(np.float64), imu.mag_ref.cpu().numpy().T.astype(np.float64)
    # oracle: pure RK4 chain and per-sample getQuarternion(acc, mag, .5, .5)  (main_file.py:40)
    for n in (0, 100, 255):
        q = np.asarray([1.0, 0.0, 0.0, 0.0])
        w = O.OracleWahba(a0[n], m0[n])
        for t in range(T):
            q = O.rk4(q, 10 ** 7, S[t, 0:3, n])
            assert O.quat_angle(gyro[t, n].cpu().numpy(), q) < 1e-5
            yw = w.quaternion(S[t, 3:6, n], S[t, 6:9, n], 0.5, 0.5)
            got = wah[t, n].cpu().numpy()
            assert O.quat_angle(got, yw) < 2e-6 and np.dot(got, yw) > 0       # same raw sign convention
    assert torch.equal(gstate.t().contiguous(), gyro[-1])
    # the Wahba-only track runs packed (two filters per thread) for even N and scalar otherwise: same bits; weights from the
    # accelerometer and the Wahba-only call go through the same kernels
    for kw in (dict(), dict(weights_from_acc=True)):
        _, w_even, _ = B.tracks(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, want_gyro=False, **kw)
        _, w_odd, _ = B.tracks(imu.streams[:, :, :255].contiguous(), imu.acc_ref[:, :255].contiguous(), imu.mag_ref[:, :255].contiguous(),
                               dt=imu.dt, want_gyro=False, **kw)
        assert torch.equal(w_even[:, :255], w_odd)
        if not kw:
            assert torch.equal(w_even, wah)
    # chunked gyro track carries its state
    g1, _, st = B.tracks(imu.streams[:77].contiguous(), imu.acc_ref, imu.mag_ref, dt=imu.dt, want_wahba=False)
    g2, _, _ = B.tracks(imu.streams[77:].contiguous(), imu.acc_ref, imu.mag_ref, dt=imu.dt, want_wahba=False, gyro_state=st)
    assert torch.equal(torch.cat([g1, g2]), gyro)
    # RPY of a stored trajectory
    rpy = B.traj2rpy(wah)
    ref = np.stack([O.quat_to_rpy_deg(wah[5, n].cpu().numpy().astype(np.float64)) for n in range(8)])
    # float32 streaming form: asin amplifies the float32 rounding of its argument by 1/cos(pitch)
    sp = np.abs(np.sin(np.radians(ref[:, 1])))
    tol = 1e-4 * np.maximum(1.0, 0.25 / np.sqrt(np.maximum(1.0 - sp * sp, 1e-12)))
    assert (np.abs(rpy[5, :8].cpu().numpy() - ref) <= tol[:, None]).all()
    # tuning objective of a small (Q,R) grid against ground truth, no trajectory stored
    truth = imu.q_true.permute(0, 2, 1).to(torch.float32).contiguous()           # [T, Ns, 4]
    grid = [(q, r) for q in (0.01, 1.0, 100.0) for r in (0.01, 0.1, 10.0)]
    G = len(grid)
    q_t = _dev(np.repeat([q for q, _ in grid], Ns), cuda)
    r_t = _dev(np.repeat([r for _, r in grid], Ns), cuda)
    st, traj, _ = B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, q=q_t, r=r_t, n_filters=G * Ns, truth=truth,
                           store_trajectory=True)
    xt, qt = traj.reshape(T, G, Ns, 4).double(), truth.double()[:, None]
    d = (xt * qt).sum(-1)
    want = ((xt * xt).sum(-1) * (qt * qt).sum(-1) - d * d).sum(0).reshape(-1)     # |X|^2|q|^2 - (X.q)^2 = sin^2
    torch.testing.assert_close(st.loss.double(), want, rtol=1e-3, atol=1e-9)
    # same loss without storing the trajectory, accumulated over two time chunks
    st2 = B.ReplayState.initial(G * Ns, cuda, r=r_t)
    loss = None
    for t0, t1 in ((0, 90), (90, T)):
        B.replay(imu.streams[t0:t1].contiguous(), imu.acc_ref, imu.mag_ref, dt=imu.dt, q=q_t, r=r_t, n_filters=G * Ns,
                 state=st2, truth=truth[t0:t1].contiguous(), loss=loss, precise_state=True)
        loss = st2.loss
    torch.testing.assert_close(loss, st.loss, rtol=1e-5, atol=1e-9)
    surface = loss.reshape(G, Ns).mean(1)
    assert torch.isfinite(surface).all() and surface.min() > 0


def test_preprocess_and_log_replay(cuda):
    """raw-sensor pre-processing kernel vs its float64 restatement, and a replay driven
    from a log in the reference's text format."""
    import os
    from poseestimationkf_b200 import logio
    rng = np.random.default_rng(3)
    T, N = 60, 300
    y1 = rng.normal(size=(T, 6, N)); y2 = y1 + 0.05 * rng.normal(size=(T, 6, N))
    t1 = rng.integers(0, 10 ** 6, (T, 2, N)); span = rng.integers(5 * 10 ** 6, 2 * 10 ** 7, (T, 2, N))
    t2 = t1 + span; t3 = t1 + (span * rng.uniform(0.05, 0.95, (T, 2, N))).astype(np.int64)
    gyro = rng.normal(size=(T, 3, N))
    tspan = np.stack([(t2 - t1)[:, 0], (t3 - t1)[:, 0], (t2 - t1)[:, 1], (t3 - t1)[:, 1]], axis=1) * 1e-9
    out, _ = B.preprocess(_dev(gyro, cuda), _dev(y1, cuda), _dev(y2, cuda), _dev(tspan, cuda))
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    for s, sl in enumerate((slice(0, 3), slice(3, 6))):
        ref = O.interpolate_normalise(f32(y1)[:, sl].transpose(0, 2, 1), f32(y2)[:, sl].transpose(0, 2, 1),
                                      t1[:, s], t2[:, s], t3[:, s])                     # [T, N, 3]
        got = out[:, 3 + 3 * s:6 + 3 * s].cpu().numpy().transpose(0, 2, 1)
        np.testing.assert_allclose(got, ref, atol=3e-6)
    np.testing.assert_array_equal(out[:, 0:3].cpu().numpy(), gyro.astype(np.float32))
    # with the low-pass stage: equals the stand-alone low-pass operator applied to the plain output
    out_lp, st = B.preprocess(_dev(gyro, cuda), _dev(y1, cuda), _dev(y2, cuda), _dev(tspan, cuda), lpf_alpha_acc=0.1,
                              lpf_alpha_mag=0.1)
    acc_lp, _ = B.lowpass(out[:, 3:6].contiguous(), 0.1)
    torch.testing.assert_close(out_lp[:, 3:6], acc_lp, rtol=1e-6, atol=1e-7)
    # replay from the text log: X_k column of the log (6 decimals) is the reference's own result
    d = logio.read_log(os.path.join(os.path.dirname(__file__), "golden", "sample_log.txt"))
    streams, acc_ref, mag_ref, dt = d.to_streams(device=cuda)
    _, traj, _ = B.replay(streams, acc_ref, mag_ref, dt=dt, q=1.0, r=0.1, store_trajectory=True)
    logged = np.asarray(d.quart_xk[1:])
    assert np.abs(traj[:, 0].cpu().numpy() - logged).max() < 2e-5        # log precision is 1e-6 per component;
    # inputs were themselves rounded to 6 decimals by the writer, hence the looser bound


def test_preprocess_and_initial_values_vs_reference_cpp_golden(cuda):
    """SURVEY 8f-2 / f-2b pinned: the pre-processing and initial-value kernels against fixtures frozen from the
    reference's OWN C++ (Parser.cpp:221-228,259-267 and InitialValues.cpp, executed through oracle/_ref by
    tests/golden/make_golden_cpp.py)."""
    import os
    from tests.conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "preprocess_ref.npz"))
    M = g["y1"].shape[0]
    H = M // 2                                          # first half accelerometer-like, second half magnetometer-like cases
    prev = np.concatenate([g["y1"][:H].T, g["y1"][H:].T])[None]          # [1, 6, H]
    nxt = np.concatenate([g["y2"][:H].T, g["y2"][H:].T])[None]
    d21, d31 = (g["t2"] - g["t1"]) * 1e-9, (g["t3"] - g["t1"]) * 1e-9    # the kernel takes the two time spans in seconds
    tspan = np.stack([d21[:H], d31[:H], d21[H:], d31[H:]])[None]
    gyro = np.zeros((1, 3, H))
    out, _ = B.preprocess(_dev(gyro, cuda), _dev(prev, cuda), _dev(nxt, cuda), _dev(tspan, cuda))
    got = out[0].cpu().numpy()
    np.testing.assert_allclose(got[3:6].T, g["normalised"][:H], atol=2e-6)
    np.testing.assert_allclose(got[6:9].T, g["normalised"][H:], atol=2e-6)
    chord = np.linalg.norm(np.concatenate([got[3:6].T, got[6:9].T]) - g["normalised"], axis=1)     # = the angle between the
    assert chord.max() < 3e-6            # unit vectors (arccos of their dot product would amplify the float32 rounding of 1)
    h = np.load(os.path.join(GOLDEN, "initial_values_ref.npz"))
    x = np.ascontiguousarray(h["samples"].transpose(1, 2, 0))             # [K, 3, N]
    mean, var = B.initial_values(_dev(x, cuda), normalize=False, want_variance=True)
    scale = np.abs(h["avg"]).max(axis=1)
    assert (np.abs(mean.cpu().numpy().T - h["avg"]) / scale[:, None]).max() < 1e-7      # float32 rounding of the output
    np.testing.assert_allclose(var.cpu().numpy().T, h["var"], rtol=2e-7)
    unit, _ = B.initial_values(_dev(x, cuda), normalize=True)
    np.testing.assert_allclose(unit.cpu().numpy().T, h["avg_unit"], atol=1e-7)


def test_wahba_negative_weights_on_device(cuda):
    rng = np.random.default_rng(5)
    M = 1000

    def unit(v):
        return v / np.linalg.norm(v, axis=-1, keepdims=True)
    arrs = [unit(rng.normal(size=(M, 3))).astype(np.float32) for _ in range(4)]
    ka = (rng.uniform(0.1, 2.0, M) * rng.choice([-1.0, 1.0], M)).astype(np.float32)
    km = (rng.uniform(0.1, 2.0, M) * rng.choice([-1.0, 1.0], M)).astype(np.float32)
    a64 = [a.astype(np.float64) for a in arrs]
    Rref = O.wahba_rotation_batched(*a64, ka.astype(np.float64), km.astype(np.float64))
    Bm = (ka.astype(np.float64)[:, None, None] * a64[0][:, :, None] * a64[2][:, None, :]
          + km.astype(np.float64)[:, None, None] * a64[1][:, :, None] * a64[3][:, None, :])
    sv = np.linalg.svd(Bm, compute_uv=False)
    ok = sv[:, 1] > 0.05 * sv[:, 0]
    for algo in ("qr2", "jacobi"):
        R, _ = B.wahba(*[_dev(a.T, cuda) for a in arrs], k_acc=_dev(ka, cuda), k_mag=_dev(km, cuda), want_rotation=True,
                       want_quaternion=False, algo=algo)
        err = np.abs(R.cpu().numpy().T.reshape(-1, 3, 3) - Rref).max(axis=(1, 2))
        assert err[ok].max() < 2e-5, (algo, err[ok].max())


def test_packed_wahba_kernel_bitwise_equals_scalar(golden_wahba, cuda):
    """posekf_wahba_f32 runs two solves per thread (f32x2) when N is even; an odd N takes the scalar
    kernel.  Same operations per solve, so dropping the last element must not change the others."""
    g = golden_wahba
    args = [_dev(g[k].T, cuda) for k in ("acc_ref", "mag_ref", "acc", "mag")]
    ka, km = _dev(g["refw_ka"], cuda), _dev(g["refw_km"], cuda)
    R2, q2 = B.wahba(*args, k_acc=ka, k_mag=km, want_rotation=True)                          # N = 1500: packed
    odd = [a[:, :1499].contiguous() for a in args]
    R1, q1 = B.wahba(*odd, k_acc=ka[:1499].contiguous(), k_mag=km[:1499].contiguous(), want_rotation=True)   # scalar
    assert torch.equal(R2[:, :1499], R1) and torch.equal(q2[:, :1499], q1)
    # shared reference pair + weights taken from the accelerometer, and the NaN-at-identity rule in a lane
    ra, rm = args[0][:, 0].contiguous(), args[1][:, 0].contiguous()
    acc = torch.stack([ra, args[2][:, 1]], dim=1).contiguous()
    mag = torch.stack([rm, args[3][:, 1]], dim=1).contiguous()
    _, qa = B.wahba(ra, rm, acc, mag, weights_from_acc=True)
    _, qb = B.wahba(ra, rm, acc[:, 1:].contiguous(), mag[:, 1:].contiguous(), weights_from_acc=True)
    assert torch.equal(qa[:, 1:], qb)


def test_initial_values(cuda):
    # SRV/InitialValues.cpp: mean of the first 100 samples, unbiased variance; Parser.cpp:46-49 normalises the mean
    rng = np.random.default_rng(2)
    x = (rng.normal(size=(100, 3, 500)) * 0.05 + rng.normal(size=(1, 3, 500))).astype(np.float32)
    mean, var = B.initial_values(_dev(x, cuda), normalize=True, want_variance=True)
    m = x.astype(np.float64).sum(axis=0) / 100
    v = ((x.astype(np.float64) - m) ** 2).sum(axis=0) / 99
    np.testing.assert_allclose(mean.cpu().numpy(), m / np.sqrt((m * m).sum(axis=0, keepdims=True)), atol=2e-6)
    np.testing.assert_allclose(var.cpu().numpy(), v, rtol=2e-4)
    raw, none = B.initial_values(_dev(x, cuda), normalize=False)
    np.testing.assert_allclose(raw.cpu().numpy(), m, atol=2e-6)
    assert none is None


def test_raw_sensor_pipeline_end_to_end(cuda):
    """The online pipeline of the C++ server as one device chain: first-100-sample means -> acc_0/mag_0
    (InitialValues / Parser.cpp:46-49), interpolation + normalisation + alpha=0.1 low-pass per gyro sample
    (Parser.cpp:229-242, KalmanFilter.cpp:279-303), then the filter -- against the float64 restatements."""
    from poseestimationkf_b200.synth import make_imu
    N, T, K = 300, 80, 100
    imu = make_imu(N, T + K, seed=9, sigma=0.01, device=cuda)
    S = imu.streams                                                       # used as "raw" unit-ish samples
    scale_a, scale_m = 9.81, 47.0                                         # raw sensors are not unit vectors
    acc_raw, mag_raw = S[:, 3:6] * scale_a, S[:, 6:9] * scale_m
    acc0, _ = B.initial_values(acc_raw[:K].contiguous(), normalize=True)
    mag0, _ = B.initial_values(mag_raw[:K].contiguous(), normalize=True)
    # bracketing samples: previous = sample t-1, next = sample t, gyro timestamp 30 % into the interval
    prev = torch.cat([acc_raw[K - 1:-1], mag_raw[K - 1:-1]], dim=1).contiguous()
    nxt = torch.cat([acc_raw[K:], mag_raw[K:]], dim=1).contiguous()
    tspan = torch.empty((T, 4, N), device=cuda)
    tspan[:, 0::2] = 0.01
    tspan[:, 1::2] = 0.003
    gyro = S[K:, 0:3].contiguous()
    streams, _ = B.preprocess(gyro, prev, nxt, tspan, lpf_alpha_acc=0.1, lpf_alpha_mag=0.1)
    _, traj, _ = B.replay(streams, acc0, mag0, dt=0.01, q=1.0, r=0.1, store_trajectory=True)
    # float64 chain
    a0 = acc_raw[:K].double().cpu().numpy().sum(0) / K; a0 = a0 / np.sqrt((a0 * a0).sum(0, keepdims=True))
    m0 = mag_raw[:K].double().cpu().numpy().sum(0) / K; m0 = m0 / np.sqrt((m0 * m0).sum(0, keepdims=True))
    np.testing.assert_allclose(acc0.cpu().numpy(), a0, atol=2e-6)
    P, Nx = prev.double().cpu().numpy(), nxt.double().cpu().numpy()
    ts = tspan.double().cpu().numpy()          # the float32 values the kernel saw
    acc_i = O.interpolate_normalise(P[:, 0:3].transpose(0, 2, 1), Nx[:, 0:3].transpose(0, 2, 1), 0.0, ts[:, 0], ts[:, 1])
    mag_i = O.interpolate_normalise(P[:, 3:6].transpose(0, 2, 1), Nx[:, 3:6].transpose(0, 2, 1), 0.0, ts[:, 2], ts[:, 3])
    Sf = np.empty((T, 9, N))
    Sf[:, 0:3] = gyro.double().cpu().numpy()
    for n in range(N):
        Sf[:, 3:6, n] = O.lowpass_scalar(acc_i[:, n], 0.1)
        Sf[:, 6:9, n] = O.lowpass_scalar(mag_i[:, n], 0.1)
    np.testing.assert_allclose(streams.cpu().numpy(), Sf, atol=3e-6)
    ref = O.replay_batched(np.full(T, 1e7), Sf[:, 0:3], Sf[:, 3:6], Sf[:, 6:9], acc0.double().cpu().numpy().T,
                           mag0.double().cpu().numpy().T, 1.0, 0.1)
    # the oracle consumes the float64 pre-processing output, the kernel its float32 one: the comparison
    # includes that input rounding (amplified while the low-pass state is still small)
    assert O.quat_angle(traj.cpu().numpy(), ref["X"])[5:].max() < 2e-5


def test_size_independent_properties_of_the_operators(cuda):
    """Properties that hold for any input size (checked at 1 Mi elements, far beyond what the oracle loops
    over): Wahba output is a proper rotation and is equivariant under a rotation of the reference frame,
    its quaternion round-trips, the low-pass is linear, RK4 preserves the norm."""
    g = torch.Generator(device=cuda); g.manual_seed(0)
    M = 1 << 20

    def unit(v):
        return v / torch.linalg.vector_norm(v, dim=0, keepdim=True)
    ra, rm, a, m = (unit(torch.randn((3, M), generator=g, device=cuda)) for _ in range(4))
    R, q = B.wahba(ra, rm, a, m, weights_from_acc=True, want_rotation=True)
    Rm = R.t().reshape(M, 3, 3)
    ok = (torch.linalg.cross(ra, rm, dim=0).norm(dim=0) > 0.2) & (torch.linalg.cross(a, m, dim=0).norm(dim=0) > 0.2) \
        & (a[2].abs() > 0.02) & (a[2].abs() < 0.98)
    eye = torch.eye(3, device=cuda)
    assert ((Rm @ Rm.transpose(1, 2) - eye).abs().amax(dim=(1, 2))[ok] < 2e-5).all()          # orthonormal
    assert ((torch.linalg.det(Rm) - 1).abs()[ok] < 2e-5).all()                                  # proper
    # quaternion of R reproduces R:  R(q) == R
    w, x, y, z = q
    Rq = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                      2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                      2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])
    assert ((Rq - R).abs().amax(dim=0)[ok] < 2e-5).all()
    # equivariance: rotating both reference vectors by Q rotates the solution by Q
    c, s = 0.6, 0.8
    Q = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], device=cuda)
    R2, _ = B.wahba((Q @ ra).contiguous(), (Q @ rm).contiguous(), a, m, weights_from_acc=True, want_rotation=True)
    assert (((Q @ Rm) - R2.t().reshape(M, 3, 3)).abs().amax(dim=(1, 2))[ok] < 3e-5).all()
    # low-pass is linear:  L(a x + b y) = a L(x) + b L(y)
    x1, x2 = torch.randn((50, 3, 4096), generator=g, device=cuda), torch.randn((50, 3, 4096), generator=g, device=cuda)
    l12, _ = B.lowpass(2.0 * x1 - 0.5 * x2, 0.1)
    l1, _ = B.lowpass(x1, 0.1); l2, _ = B.lowpass(x2, 0.1)
    torch.testing.assert_close(l12, 2.0 * l1 - 0.5 * l2, rtol=1e-4, atol=1e-5)
    # RK4 + normalise keeps unit norm; zero rate is the identity
    q0 = unit(torch.randn((4, M), generator=g, device=cuda))
    w3 = torch.randn((3, M), generator=g, device=cuda)
    q1 = B.rk4(q0, 0.01, w3)
    assert ((q1.norm(dim=0) - 1).abs() < 3e-7).all()
    assert (B.rk4(q0, 0.01, torch.zeros_like(w3)) - q0).abs().max() < 5e-7      # only the renormalisation of q0
