#!/usr/bin/env python
"""bench.py -- EKF filter-steps/s of the batched replay on N B200s (contract in the task prompt).

Workload (BASELINE.json configs[1]): 1 Mi independent filters x 1000 steps, float32, per GPU
(weak scaling: every rank replays its own 1 Mi filters; filters are independent, so there is no
collective on the path -- the only communication is the MAX-over-ranks of the timing).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference ...                          the reference's CPU path (oracle port)

A "step" of the contract = one full pass of the hot path over the resident workload (one replay of
all T timesteps of all filters).  `value` = filter-steps/s with inputs resident in HBM;
`e2e` = the same metric through posekf_replay_host_f32 with pinned HOST buffers (H2D of the
stream and D2H of the final state inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ekf_filter_steps_per_s"
UNIT = "filter-steps/s"
# algorithmic work per filter-step (DESIGN.md "Roofline"): bytes streamed and flops executed
BYTES_PER_STEP = 36            # 9 float32 inputs (final-state-only replay)
# flops of the algorithm the kernel EXECUTES (FMA = 2, mul/add/rcp/rsqrt = 1; compares and selects
# not counted), stage by stage in DESIGN.md "Roofline"; the SASS FFMA/FMUL/FADD/MUFU census of the
# loop body gives the same number.  (SURVEY.md's 1570 is the un-restructured reference algorithm.)
# qr2: dynamic opcode census of the shipped packed kernel (profiles/r01_replay_packed_opcode_census.json):
# 121.0 FMA + 79.4 MUL + 39.1 ADD lane operations + 8 MUFU per filter-step = 368.6 flops, 239.5 FP32 lane operations
# (+ 10 FSEL, which also issue to the FP32 pipe on sm_100).
FLOPS = {"qr2": 368.6, "jacobi": 1292}
FP32_LANE_OPS = {"qr2": 249.5}    # FP32-pipe lane operations per filter-step (FMA, MUL and ADD each occupy one lane-cycle)
FLOPS_COMPENSATED_EXTRA = 32      # two-sum folding of the state (ncu: 536 + 10 MUFU flops per filter-step)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filters", type=int, default=1 << 20, help="filters per GPU")
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--wahba", default="qr2", choices=["qr2", "jacobi"])
    ap.add_argument("--staging", default="auto", choices=["auto", "ldg", "tma", "tma_packed"])
    ap.add_argument("--e2e-timesteps", type=int, default=250, help="timesteps of the host-buffer (e2e) replay")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the CPUs local to its GPU")
    ap.add_argument("--no-variants", action="store_true", help="skip the context measurements of the other kernels")
    ap.add_argument("--precise-state", action="store_true",
                    help="compensated two-float state (needed for R >> Q sweeps, not for the Q=1, R=0.1 headline config)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's scalar port (same numpy call sequence as the reference classes),
# one independent trajectory per worker process, all host cores.
# ------------------------------------------------------------------------------------------------
_BARRIER = None


def _cpu_init(barrier):
    global _BARRIER
    _BARRIER = barrier
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"


def _cpu_worker(args):
    """One independent trajectory through the oracle's scalar port.  Input synthesis is untimed; all
    workers start the timed replay together (barrier) so the rate is a genuine all-cores figure."""
    seed, n_steps, use_barrier = args
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import ekf_oracle as O
    from poseestimationkf_b200.synth import make_imu
    imu = make_imu(1, n_steps, seed=seed, sigma=0.01)
    S = imu.streams.numpy().astype(np.float64)
    t_ns = np.arange(n_steps + 1, dtype=np.int64) * 10 ** 7
    a0, m0 = imu.acc_ref[:, 0].numpy().astype(np.float64), imu.mag_ref[:, 0].numpy().astype(np.float64)
    if use_barrier and _BARRIER is not None:
        _BARRIER.wait()
    t0 = time.perf_counter()
    X, _ = O.replay_scalar(t_ns, S[:, 0:3, 0], S[:, 3:6, 0], S[:, 6:9, 0], a0, m0, 1.0, 0.1)
    return time.perf_counter() - t0, float(X[-1, 0])


def cpu_baseline(target_seconds: float, cores: int | None = None):
    """Times the reference path (oracle port, float64 numpy) on `cores` processes; returns the
    cpu_baseline object.  Sample: one independent 100 Hz trajectory per core."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    # calibrate on one short run in this process
    dt1, _ = _cpu_worker((0, 300, False))
    per_step = dt1 / 300
    n_steps = int(max(500, min(20000, target_seconds / per_step)))
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(cores)
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(barrier,)) as pool:
        pool.map(_cpu_worker, [(i, 50, False) for i in range(cores)], chunksize=1)          # warm the workers (imports)
        res = pool.map(_cpu_worker, [(100 + i, n_steps, True) for i in range(cores)], chunksize=1)
    wall = max(r[0] for r in res)
    single = 1.0 / per_step
    return {"value": cores * n_steps / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} independent synthetic 100 Hz trajectories x {n_steps} steps, Q=1 R=0.1, float64 numpy "
                      f"oracle port of the reference classes (one process per core, started together); "
                      f"single-core rate {single:.0f} steps/s",
            "single_core_value": single, "seconds": wall}


def cpu_baseline_compiled(n_filters: int = 16384, n_steps: int = 500):
    """The same path as an optimised CPU program would run it: oracle/ekf_oracle.c (float64, dense 4x4
    algebra + Jacobi SVD as the reference's numpy calls do, gcc -O2), one pthread per host core.
    Reported beside `cpu_baseline` (the reference's own Python implementation) for context."""
    import numpy as np
    from oracle import c_oracle as CO
    from poseestimationkf_b200.synth import make_imu
    cores = os.cpu_count() or 1
    imu = make_imu(n_filters, n_steps, seed=5, sigma=0.01)
    S = imu.streams.numpy()
    ar, mr = imu.acc_ref.numpy(), imu.mag_ref.numpy()
    CO.replay(S[:50], 1e7, ar, mr, 1.0, 0.1, store=False, flips=False, threads=cores)
    t0 = time.perf_counter()
    CO.replay(S, 1e7, ar, mr, 1.0, 0.1, store=False, flips=False, threads=cores)
    wall = time.perf_counter() - t0
    return {"value": n_filters * n_steps / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_filters} filters x {n_steps} steps, float64 C restatement (oracle/ekf_oracle.c), {cores} pthreads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(max(5.0, args.cpu_seconds))
    # K "steps" of the contract = K repetitions of the bounded sample; report the mean rate
    vals = [base["value"]]
    for _ in range(max(0, min(args.steps, 3) - 1)):
        vals.append(cpu_baseline(max(5.0, args.cpu_seconds))["value"])
    v = sum(vals) / len(vals)
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * base["seconds"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"batched EKF replay: {args.filters} independent filters x {args.timesteps} steps per GPU, Q=1, R=0.1, "
                                   "dt=0.01 (BASELINE.json configs[1]); reference arm = the reference's Python EKF (oracle port, "
                                   "float64 numpy) on all host cores, each step a bounded sample of that workload (one independent "
                                   "trajectory per core)",
                       "filters_per_gpu": args.filters, "timesteps": args.timesteps, "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = sorted(int(float(r[0])) for r in rows if r and r[0].replace(".", "").isdigit())
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active") and name not in reasons:
                    reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(float(rows[0][1])) if rows else None,
                "power_w_max": max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from poseestimationkf_b200 import batched as B
    from poseestimationkf_b200 import sharding as SH
    from poseestimationkf_b200.synth import make_imu

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: keep this rank's pinned staging memory (e2e) on the socket its GPU hangs off
    all_cpus = os.sched_getaffinity(0)
    affinity = None if args.no_numa_bind else SH.bind_host_to_gpu(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, T = args.filters, args.timesteps
    K, W = args.steps, max(args.warmup, 3)

    # --- synthetic workload, resident in HBM: [T, 9, N] float32 (36 GB at 1 Mi x 1000) ---------
    # 16 Ki distinct trajectories are generated (float64 ground truth on the device), replicated
    # along the filter axis, and every replica gets its own additive sensor noise and (Q,R) so that
    # no two filters do the same arithmetic.
    base_n = min(N, 1 << 14)
    imu = make_imu(base_n, T, seed=1000 + rank, sigma=0.0, device=dev)
    reps = (N + base_n - 1) // base_n
    streams = torch.empty((T, 9, N), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    for t0 in range(0, T, 50):
        blk = imu.streams[t0:t0 + 50].repeat(1, 1, reps)[:, :, :N]
        blk = blk + 0.01 * torch.randn(blk.shape, generator=g, device=dev)
        for sl in (slice(3, 6), slice(6, 9)):
            blk[:, sl] = blk[:, sl] / torch.linalg.vector_norm(blk[:, sl], dim=1, keepdim=True)
        streams[t0:t0 + 50] = blk
        del blk
    acc_ref = imu.acc_ref.repeat(1, reps)[:, :N].contiguous()
    mag_ref = imu.mag_ref.repeat(1, reps)[:, :N].contiguous()
    q = torch.full((N,), 1.0, device=dev)
    r = torch.full((N,), 0.1, device=dev)
    del imu
    torch.cuda.synchronize()

    def one_pass(state):
        B.replay(streams, acc_ref, mag_ref, dt=0.01, q=q, r=r, state=state, wahba=args.wahba, staging=args.staging,
                 precise_state=args.precise_state)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp32_peak, _ = B.fp32_peak_tflops(local)
    torch.cuda.profiler.start()        # ncu --profile-from-start off: list warm-up + timed launches only
    for w in range(W):
        one_pass(B.ReplayState.initial(N, dev, r=0.1))
    barrier()

    # --- timed region: exactly K passes; per-launch CUDA events on the launching stream ---------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    fresh = [B.ReplayState.initial(N, dev, r=0.1) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record()
    for k in range(K):
        evs[k][0].record()
        one_pass(fresh[k])
        evs[k][1].record()
    e_all1.record()
    barrier()
    t_wall1 = time.perf_counter()
    torch.cuda.profiler.stop()
    total_ms = e_all0.elapsed_time(e_all1)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    if world > 1:
        total_ms = SH.max_over_ranks(total_ms, dev)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    ms_per_step = total_ms / K
    value = world * N * T / (ms_per_step * 1e-3)
    avg_kernel_ms = sum(kernel_ms) / len(kernel_ms)
    steps_per_s_kernel = N * T / (avg_kernel_ms * 1e-3)

    # --- context: the other kernels on the first 250 timesteps of the same resident workload (not the headline) ----
    variants = {}
    if rank == 0 and not args.no_variants:
        Tv = min(250, T)
        sub = streams[:Tv]
        traj_buf = torch.empty((Tv, N, 4), dtype=torch.float32, device=dev)

        def rate(**kw):
            best = 1e30
            for i in range(3):
                st = B.ReplayState.initial(N, dev, r=0.1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                B.replay(sub, acc_ref, mag_ref, dt=0.01, q=q, r=r, state=st, **kw)
                e1.record()
                torch.cuda.synchronize()
                if i:
                    best = min(best, e0.elapsed_time(e1))
            return N * Tv / (best * 1e-3)

        variants = {
            "packed_tma (default)": rate(precise_state=False, staging="tma_packed"),
            "scalar_tma": rate(precise_state=False, staging="tma"),
            "scalar_ldg": rate(precise_state=False, staging="ldg"),
            "packed_tma + precise variant (sweeps)": rate(precise_state=True, staging="tma_packed"),
            "packed_tma + trajectory stored [T,N,4]": rate(precise_state=False, staging="tma_packed", out_traj=traj_buf),
            "scalar_tma, Jacobi SVD Wahba (north_star literal)": rate(precise_state=False, staging="tma", wahba="jacobi"),
        }
        del traj_buf
        variants = {k: round(v / 1e9, 2) for k, v in variants.items()}

    # --- e2e: host buffers through the C ABI (H2D + kernel + D2H inside the timed region) --------
    e2e = None
    if not args.no_e2e:
        Te = min(args.e2e_timesteps, T)
        host = torch.empty((Te, 9, N), dtype=torch.float32, pin_memory=True)
        host.copy_(streams[:Te])
        ar_h, mr_h = acc_ref.cpu().pin_memory(), mag_ref.cpu().pin_memory()
        q_h, r_h = q.cpu().pin_memory(), r.cpu().pin_memory()
        # host-link reference: one plain pinned H2D copy of the same buffer (what the link can do)
        scratch = torch.empty_like(streams[:Te])
        scratch.copy_(host, non_blocking=True); torch.cuda.synchronize()
        tl0 = time.perf_counter()
        scratch.copy_(host, non_blocking=True); torch.cuda.synchronize()
        link_gbs = host.numel() * 4 / (time.perf_counter() - tl0) / 1e9
        del scratch
        ws = B.HostWorkspace(N, device=local)
        # result buffers pinned once (a fresh 58 MB pinned allocation per call costs ~10 ms of page pinning)
        x_h = torch.empty((4, N), dtype=torch.float32, pin_memory=True)
        p_h = torch.empty((10, N), dtype=torch.float32, pin_memory=True)
        B.replay_host(host, ar_h, mr_h, dt=0.01, q=q_h, r=r_h, wahba=args.wahba, device=local, workspace=ws,
                      precise_state=args.precise_state, out_x=x_h, out_p=p_h)   # warm-up
        barrier()
        reps_e = 3
        t0 = time.perf_counter()
        for _ in range(reps_e):
            B.replay_host(host, ar_h, mr_h, dt=0.01, q=q_h, r=r_h, wahba=args.wahba, device=local,
                          workspace=ws, precise_state=args.precise_state, out_x=x_h, out_p=p_h)
        barrier()
        dt_e = (time.perf_counter() - t0) / reps_e
        if world > 1:
            dt_e = SH.max_over_ranks(dt_e, dev)
        h2d = host.numel() * 4 + (ar_h.numel() + mr_h.numel() + q_h.numel() + r_h.numel()) * 4
        d2h = (x_h.numel() + p_h.numel()) * 4
        e2e = {"value": world * N * Te / dt_e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "timesteps": Te, "seconds_per_pass": dt_e,
               "note": "posekf_replay_host_f32: pinned host [T,9,N] stream -> double-buffered H2D chunks -> kernel -> "
                       "D2H of final X,P; bound by the host link (36 B per filter-step)",
               "h2d_gbs": h2d / dt_e / 1e9, "host_link_plain_copy_gbs": link_gbs, "host_affinity": affinity}
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    flops = FLOPS[args.wahba] + (FLOPS_COMPENSATED_EXTRA if args.precise_state else 0)
    packed = args.staging in ("auto", "tma_packed") and args.wahba == "qr2" and N % 4 == 0
    ach_tf = steps_per_s_kernel * flops / 1e12
    ach_gbs = steps_per_s_kernel * BYTES_PER_STEP / 1e9
    # the binding roof is the slower of FP32 issue and HBM streaming (north_star); report both
    t_fp32 = flops / (fp32_peak * 1e12)
    t_hbm = BYTES_PER_STEP / (hbm_peak * 1e9)
    bound = "fp32" if t_fp32 >= t_hbm else "hbm"
    roofline = {
        "bound": bound,
        "achieved": ach_tf if bound == "fp32" else ach_gbs,
        "peak": fp32_peak if bound == "fp32" else hbm_peak,
        "unit": "TFLOP/s" if bound == "fp32" else "GB/s",
        "frac": (ach_tf / fp32_peak) if bound == "fp32" else (ach_gbs / hbm_peak),
        "traffic": None,
        "kernel": ("replay_tma2_kernel (packed f32x2, two filters per thread)" if packed
                   else f"replay_{'tma' if args.staging != 'ldg' else 'ldg'}_kernel<{args.wahba}>"),
        "kernel_ms": avg_kernel_ms,
        "algorithmic_flops_per_filter_step": flops,
        "algorithmic_bytes_per_filter_step": BYTES_PER_STEP,
        "fp32": {"achieved_tflops": ach_tf, "peak_tflops": fp32_peak, "frac": ach_tf / fp32_peak,
                 "peak_source": "FFMA probe kernel measured in this run (posekf_fp32_peak_tflops)"},
        "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": hbm_peak, "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
        "roofline_steps_per_s_per_gpu": 1.0 / max(t_fp32, t_hbm),
        # the flop count has 1.6 flops per FP32 instruction (MUL and ADD carry one), so a 100 % busy pipe is < 100 % of
        # the FFMA peak; the share of the pipe's lane-cycles (148 SMs x 128 lanes x SM clock) the kernel fills:
        "fp32_pipe_lane_cycles_frac": (steps_per_s_kernel * FP32_LANE_OPS[args.wahba]
                                       / (torch.cuda.get_device_properties(dev).multi_processor_count * 128
                                          * (clocks or {}).get("sm_mhz", 0) * 1e6)
                                       if args.wahba in FP32_LANE_OPS and not args.precise_state and (clocks or {}).get("sm_mhz") else None),
        "frac_of_roofline": steps_per_s_kernel * max(t_fp32, t_hbm),
    }
    # ncu traffic: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
    # (profiles/r01_replay_packed_ncu_full.json, taken at 200 timesteps), scaled per launch to this run's timesteps;
    # the scalar and precise kernels have older captures under profiles/history/
    try:
        name = ("history/r01_replay_tma_compensated_ncu_full.json" if args.precise_state else
                "r01_replay_packed_ncu_full.json" if packed else "history/r01_replay_tma_ncu_full.json")
        prof = json.load(open(os.path.join(ROOT, "profiles", name)))
        per_step = prof["dram_bytes_per_launch"] / prof["workload"]["filter_steps"]
        roofline["traffic"] = per_step * N * T
        roofline["traffic_source"] = f"profiles/{name}: {per_step:.2f} B per filter-step measured vs {BYTES_PER_STEP} algorithmic"
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batched EKF replay: {N} independent filters x {T} steps per GPU, Q=1, R=0.1, dt=0.01 "
                               "(BASELINE.json configs[1])",
                   "filters_per_gpu": N, "timesteps": T, "wahba": args.wahba, "staging": args.staging,
                   "store_trajectory": False, "state": "two-float compensated" if args.precise_state else "float32",
                   "parallelism": f"filter-sharded x{world}, no collective",
                   "l2_policy": f"input stream is {N * T * 36 / 1e9:.1f} GB per pass (>> 126 MB L2), streamed once"},
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": K,
        "clocks": clocks,
        "variants_gsteps_per_s_1gpu_250_timesteps": variants,
    }
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)          # the CPU arm uses every host core again
        line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        try:
            line["cpu_baseline_compiled"] = cpu_baseline_compiled()
        except Exception as exc:      # the C checker is optional infrastructure
            line["cpu_baseline_compiled"] = {"unavailable": str(exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
