"""BASELINE.json configs 3, 4 and 5 at FULL size on one B200, each checked against the compiled float64 oracle
(oracle/ekf_oracle.c, itself held to the unmodified reference's outputs by tests/test_oracle_c.py) on a sample of
at least 1024 filters.  Tolerance: 1e-5 rad quaternion angle (BASELINE.json north_star), same q/-q sign.
The workloads are the ones `bench.py --workload c3|c4|c5` times (poseestimationkf_b200/workloads.py)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import ekf_oracle as O
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200 import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_c3_tuning_sweep_full_size(cuda):
    """64x64 (Q,R) grid x 256 trajectories x 5000 steps = 5.24e9 filter-steps in one call: automatic precision per
    cell, shared measurements, loss surface on the device.  Parity of the final state at the four corners of the
    grid and at the reference's own tuning (Q=1, R=0.1): 5 x 256 = 1280 filters through the oracle."""
    w = W.build_c3(cuda)
    assert w.n_filters == 64 * 64 * 256 and w.streams.shape == (5000, 9, 256)
    st = W.run_c3(w)
    torch.cuda.synchronize()
    assert st.x_lo is not None                      # some cells needed the precise variant
    S = w.streams.cpu().numpy()
    ar, mr = w.acc_ref.cpu().numpy(), w.mag_ref.cpu().numpy()
    G = 64
    cells = [(0, G - 1), (G - 1, 0), (0, 0), (G - 1, G - 1), (30, 20)]
    assert abs(float(w.qs[30]) - 1.0) < 1e-5 and abs(float(w.rs[20]) - 0.1) < 1e-6
    for iq, ir in cells:
        ref = CO.replay(S, w.dt * 1e9, ar, mr, float(w.qs[iq]), float(w.rs[ir]), store=False, flips=False)
        got = st.x[:, w.cell(iq, ir)].t().cpu().numpy().astype(np.float64)
        ang = O.quat_angle(got, ref["X_final"])
        assert ang.max() < TOL, (iq, ir, ang.max())
        assert (np.sum(got * ref["X_final"], axis=1) > 0).all()
    surf = W.loss_surface(w, st)
    assert surf.shape == (G, G) and bool(torch.isfinite(surf).all()) and float(surf.min()) > 0
    # the tuning objective itself: recompute one cell's loss in float64 from the oracle's trajectory
    iq, ir = 30, 20
    ref = CO.replay(S[:, :, :32], w.dt * 1e9, ar[:, :32], mr[:, :32], float(w.qs[iq]), float(w.rs[ir]), store=True, flips=False)
    truth = w.truth[:, :32].cpu().numpy().astype(np.float64)
    d = np.sum(ref["X"] * truth, axis=-1)
    want = np.sum(1.0 - d * d / (np.sum(ref["X"] ** 2, -1) * np.sum(truth ** 2, -1)), axis=0)
    got = st.loss[w.cell(iq, ir)][:32].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-7)


@pytest.mark.parametrize("weights", ["half", "reference"])
def test_c4_wahba_only_full_size(cuda, weights):
    """100 M (acc, mag) pairs -> reference-signed quaternion, both weightings; 4096 evenly spaced pairs through the
    oracle: angle and sign."""
    w = W.build_c4(cuda)
    M = w.acc.shape[1]
    assert M == 100_000_000
    out = W.run_c4(w, "qr2", weights)
    torch.cuda.synchronize()
    idx = torch.arange(0, M, M // 4096, device=cuda)[:4096]
    acc, mag = w.acc[:, idx].cpu().numpy(), w.mag[:, idx].cpu().numpy()
    ka = np.abs(acc[2]).astype(np.float64) if weights == "reference" else np.full(4096, 0.5)
    km = (1 - ka) if weights == "reference" else ka
    ra = w.acc_ref.cpu().numpy()[:, None].repeat(4096, 1)
    rm = w.mag_ref.cpu().numpy()[:, None].repeat(4096, 1)
    _, qref = CO.wahba(ra, rm, acc, mag, ka, km)
    got = out[:, idx].t().cpu().numpy().astype(np.float64)
    assert O.quat_angle(got, qref).max() < TOL
    assert (np.sum(got * qref, axis=1) > 0).all()            # RotationMatrix2Quart's sign convention
    # size-independent properties over all 100 M outputs: unit norm, finite
    nrm = torch.linalg.vector_norm(out, dim=0)
    assert bool(torch.isfinite(out).all()) and float((nrm - 1).abs().max()) < 1e-6
    # the (QR-preconditioned) Jacobi SVD reaches the same answer under both weightings
    outj = W.run_c4(w, "jacobi", weights)
    assert O.quat_angle(outj[:, idx].t().cpu().numpy().astype(np.float64), qref).max() < TOL


def test_c5_long_sharded_replay_one_rank_share(cuda):
    """One rank's share of config 5 on 8 GPUs: 2 Mi filters x 2000 steps in time chunks generated on the device, state
    carried in the kernel's frame.  The first 1024 filters (= the first 1024 base trajectories) against the oracle,
    and the chunked replay against an unchunked one on a slice that fits."""
    job = W.ShardedLongReplay(cuda, rank=0, world=8)
    assert job.n_local == (1 << 24) // 8 and job.T == 2000 and job.launches_per_pass >= 8
    st = job.new_state()
    ms = job.run_pass(st)
    assert st.frame == "reference" and ms > 0
    base = job.base
    ref = CO.replay(base.streams[:, :, :1024].cpu().numpy(), base.dt * 1e9, base.acc_ref[:, :1024].cpu().numpy(),
                    base.mag_ref[:, :1024].cpu().numpy(), float(np.float32(1.0)), float(np.float32(0.1)), store=False, flips=False)
    got = st.x[:, :1024].t().cpu().numpy().astype(np.float64)
    ang = O.quat_angle(got, ref["X_final"])
    assert ang.max() < TOL, ang.max()
    assert (np.sum(got * ref["X_final"], axis=1) > 0).all()
    # every replica of a base trajectory holds the same bits (same inputs, same arithmetic) ...
    assert torch.equal(st.x[:, :job.base_n], st.x[:, job.base_n:2 * job.base_n])
    # ... and they equal ONE unchunked launch over the whole 2000 steps
    whole, _, _ = B.replay(base.streams[:, :, :4096].contiguous(), base.acc_ref[:, :4096].contiguous(),
                           base.mag_ref[:, :4096].contiguous(), dt=base.dt, q=1.0, r=0.1, precise_state=False)
    assert torch.equal(whole.x, st.x[:, :4096]) and torch.equal(whole.p, st.p[:, :4096])
    # a shard that does not start at column 0 of the base batch reads the right columns
    other = W.ShardedLongReplay(cuda, rank=3, world=5, n_filters=5 * 128 * 100 + 640, n_steps=8, base_n=1000)
    view = other.fill_chunk(0, 8)
    cols = torch.arange(other.begin, other.end, device=cuda) % other.base_n
    assert torch.equal(view, other.base.streams[:, :, cols]) and torch.equal(other.acc_ref, other.base.acc_ref[:, cols])
