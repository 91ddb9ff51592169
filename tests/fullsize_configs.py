"""Runs the other BASELINE.json configs at full size on one GPU and prints one JSON line each
(supplementary evidence for profiles/; bench.py is the contract benchmark).  Lives under tests/ because it
checks every result against the float64 oracle (only tests/, smoke() and bench.py's CPU arm may use oracle/).
    python tests/fullsize_configs.py [c3] [c4] [c5]

  C3  Q/R tuning sweep: 64x64 log-spaced (Q,R) grid x 256 trajectories x 5000 steps (1 Mi filters,
      shared 46 MB stream, on-device loss surface, compensated state)
  C4  Wahba-only: 100 M (acc, mag) pairs -> quaternion, both solvers
  C5  per-GPU share of 16 Mi filters x 2000 steps on 8 GPUs: 2 Mi filters x 2000 steps, time-chunked
      with carried state (inputs generated per chunk, only the filter kernels are timed)
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
from oracle import c_oracle as CO
from oracle import ekf_oracle as O
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200 import sharding as SH
from poseestimationkf_b200.synth import make_imu

dev = torch.device("cuda:0")
which = sys.argv[1:] or ["c3", "c4", "c5"]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


if "c3" in which:
    Ns, T, G1 = 256, 5000, 64
    imu = make_imu(Ns, T, seed=3, sigma=0.01, device=dev, keep_truth=True)
    qs = torch.logspace(-3, 3, G1, device=dev); rs = torch.logspace(-3, 3, G1, device=dev)
    q_t = qs.repeat_interleave(G1).repeat_interleave(Ns).contiguous()        # grid index g = iq*64 + ir
    r_t = rs.repeat(G1).repeat_interleave(Ns).contiguous()
    N = G1 * G1 * Ns
    truth = imu.q_true.permute(0, 2, 1).to(torch.float32).contiguous()
    out = {}
    for precise, share in ((None, True), (True, True), (True, False), (False, True), (False, False)):
        def run():
            st = B.ReplayState.initial(N, dev, r=r_t)
            B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, q=q_t, r=r_t, n_filters=N, state=st, truth=truth,
                     precise_state=precise, share_measurements=share)
            out["st"] = st
        ms = timed(run)
        st = out["st"]
        surface = (st.loss.reshape(G1, G1, Ns).mean(2) / T).cpu().numpy()
        # parity of the final state at the four corners + the reference tuning vs the compiled oracle
        S = imu.streams.cpu().numpy()
        worst = {}
        for iq, ir in ((0, G1 - 1), (G1 - 1, 0), (0, 0), (G1 - 1, G1 - 1), (G1 // 2, G1 // 3)):
            g = iq * G1 + ir
            ref = CO.replay(S, imu.dt * 1e9, imu.acc_ref.cpu().numpy(), imu.mag_ref.cpu().numpy(), float(qs[iq]), float(rs[ir]),
                            store=False, flips=False)
            got = st.x[:, g * Ns:(g + 1) * Ns].t().cpu().numpy()
            worst[f"q={float(qs[iq]):.0e},r={float(rs[ir]):.0e}"] = float(O.quat_angle(got, ref["X_final"]).max())
        best = np.unravel_index(surface.argmin(), surface.shape)
        print(json.dumps({"config": "C3 sweep 64x64 (Q,R) x 256 trajectories x 5000 steps", "precise_state": precise,
                          "share_measurements": share,
                          "filters": N, "ms": ms, "gsteps_per_s": N * T / ms / 1e6,
                          "final_state_max_angle_vs_oracle": worst,
                          "loss_surface_min_at": {"q": float(qs[best[0]]), "r": float(rs[best[1]]), "mean_sin2": float(surface.min())}}))
    del imu, truth

if "c4" in which:
    M = 100_000_000
    g = torch.Generator(device=dev); g.manual_seed(1)
    def unit(v): return v / torch.linalg.vector_norm(v, dim=0, keepdim=True)
    acc = unit(torch.randn((3, M), generator=g, device=dev)); mag = unit(torch.randn((3, M), generator=g, device=dev))
    ra = torch.tensor([0.0, 0.0, 1.0], device=dev); rm = unit(torch.tensor([[0.4], [0.0], [-0.9165]], device=dev))[:, 0].contiguous()
    qout = torch.empty((4, M), device=dev)
    for algo in ("qr2", "jacobi"):
        for wname, kw in (("half", dict(k_acc=0.5, k_mag=0.5)), ("reference", dict(weights_from_acc=True))):
            def run():
                B._lib.check(B._lib.load().posekf_wahba_f32(M, ra.data_ptr(), rm.data_ptr(), 1, acc.data_ptr(), mag.data_ptr(),
                             None, None, kw.get("k_acc", 0.0), kw.get("k_mag", 0.0), int(kw.get("weights_from_acc", False)),
                             None, qout.data_ptr(), B._lib.WAHBA[algo], 0, torch.cuda.current_stream().cuda_stream), "wahba")
            ms = timed(run)
            idx = torch.arange(0, M, M // 4096, device=dev)[:4096]
            ka = acc[2, idx].abs().double().cpu().numpy() if wname == "reference" else np.full(4096, 0.5)
            _, qref = CO.wahba(ra.cpu().numpy()[:, None].repeat(4096, 1), rm.cpu().numpy()[:, None].repeat(4096, 1),
                               acc[:, idx].cpu().numpy(), mag[:, idx].cpu().numpy(), ka, (1 - ka) if wname == "reference" else ka)
            ang = O.quat_angle(qout[:, idx].t().cpu().numpy(), qref)
            print(json.dumps({"config": "C4 Wahba-only 100M pairs", "algo": algo, "weights": wname, "ms": ms,
                              "gsolves_per_s": M / ms / 1e6, "hbm_gbs": M * 40 / ms / 1e6,
                              "max_angle_vs_oracle_4096_sample": float(ang.max()), "median": float(np.median(ang))}))
    del acc, mag, qout

if "c5" in which:
    N, T, chunk = 2 * (1 << 20), 2000, 250
    base = make_imu(1 << 14, T, seed=8, sigma=0.01, device=dev)
    reps = N // (1 << 14)
    acc_ref, mag_ref = base.acc_ref.repeat(1, reps).contiguous(), base.mag_ref.repeat(1, reps).contiguous()
    st = B.ReplayState.initial(N, dev, r=0.1)
    q = torch.full((N,), 1.0, device=dev); r = torch.full((N,), 0.1, device=dev)
    buf = torch.empty((chunk, 9, N), device=dev)
    # untimed warm-up on a throw-away state: the first launch of a kernel pays CUDA's lazy module loading (tens of ms)
    B.replay(base.streams[:4].repeat(1, 1, reps), acc_ref, mag_ref, dt=base.dt, q=q, r=r,
             state=B.ReplayState.initial(N, dev, r=0.1), precise_state=False)
    total_ms = 0.0
    for t0, t1 in SH.time_chunks(T, chunk):
        buf[: t1 - t0] = base.streams[t0:t1].repeat(1, 1, reps)             # untimed: synthetic input for this chunk
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.replay(buf[: t1 - t0], acc_ref, mag_ref, dt=base.dt, q=q, r=r, state=st, precise_state=False, keep_filter_frame=t1 < T)
        e1.record(); torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
    ref = CO.replay(base.streams[:, :, :512].cpu().numpy(), base.dt * 1e9, base.acc_ref[:, :512].cpu().numpy(),
                    base.mag_ref[:, :512].cpu().numpy(), float(np.float32(1.0)), float(np.float32(0.1)), store=False, flips=False)
    ang = O.quat_angle(st.x[:, :512].t().cpu().numpy(), ref["X_final"])
    print(json.dumps({"config": "C5 per-GPU share: 2 Mi filters x 2000 steps, 8 time chunks of 250 with carried state",
                      "kernel_ms_total": total_ms, "gsteps_per_s": N * T / total_ms / 1e6,
                      "final_state_max_angle_vs_oracle_512_filters": float(ang.max())}))
