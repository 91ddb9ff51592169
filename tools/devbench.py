"""Developer micro-bench: replay throughput by staging / Wahba algorithm (not the contract bench)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200.synth import make_imu

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--t", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variants", default="ldg:qr2,tma:qr2,ldg:jacobi,tma:jacobi")
ap.add_argument("--traj", type=int, default=0)
ap.add_argument("--ns", type=int, default=0, help="distinct trajectories (sweep layout); 0 = one per filter")
ap.add_argument("--tag", default="")
ap.add_argument("--base", type=int, default=4096, help="filters in the synthetic batch that is tiled up to --n")
ap.add_argument("--sustain", type=int, default=0,
                help="launch this many passes back to back (no idle gaps) and also report the mean of the last half: the rate "
                     "under the power cap, as the contract bench sees it")
args = ap.parse_args()

dev = torch.device("cuda:0")
print(json.dumps({"fp32_peak_tflops": B.fp32_peak_tflops(0)[0]}))
# build the stream by tiling a smaller synthetic batch (values do not matter for timing)
base = make_imu(args.base, args.t, seed=1, sigma=0.01, device=dev)
if args.ns:
    streams, acc_ref, mag_ref = (base.streams[:, :, :args.ns].contiguous(), base.acc_ref[:, :args.ns].contiguous(),
                                 base.mag_ref[:, :args.ns].contiguous())
    N = args.n
else:
    reps = args.n // args.base
    streams = base.streams.repeat(1, 1, reps).contiguous()
    acc_ref = base.acc_ref.repeat(1, reps).contiguous()
    mag_ref = base.mag_ref.repeat(1, reps).contiguous()
    N = streams.shape[2]
traj = torch.empty((args.t, N, 4), dtype=torch.float32, device=dev) if args.traj else None
for var in args.variants.split(","):
    staging, algo = var.split(":")
    times = []
    for r in range(args.reps + 1):
        st = B.ReplayState.initial(N, dev, r=0.1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.replay(streams, acc_ref, mag_ref, dt=0.01, q=1.0, r=0.1, state=st, out_traj=traj, wahba=algo, staging=staging, n_filters=N)
        e1.record()
        torch.cuda.synchronize()
        if r:
            times.append(e0.elapsed_time(e1))
    ms = min(times)
    steps = N * args.t
    sustained = None
    if args.sustain:
        sts = [B.ReplayState.initial(N, dev, r=0.1) for _ in range(args.sustain)]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.sustain)]
        torch.cuda.synchronize()
        for st_, (a, b) in zip(sts, evs):
            a.record()
            B.replay(streams, acc_ref, mag_ref, dt=0.01, q=1.0, r=0.1, state=st_, out_traj=traj, wahba=algo, staging=staging, n_filters=N)
            b.record()
        torch.cuda.synchronize()
        per = [a.elapsed_time(b) for a, b in evs]
        sustained = sum(per[len(per) // 2:]) / (len(per) - len(per) // 2)
    print(json.dumps({"tag": args.tag, "ns": args.ns, "variant": var, "N": N, "T": args.t, "traj": bool(args.traj), "ms": round(ms, 3),
                      "gsteps_per_s": round(steps / ms / 1e6, 2),
                      "sustained_gsteps_per_s": None if sustained is None else round(steps / sustained / 1e6, 2), "hbm_gbs": round(steps * (36 + 16 * bool(args.traj)) / ms / 1e6, 1),
                      "x0": st.x[:, 0].tolist()}))
