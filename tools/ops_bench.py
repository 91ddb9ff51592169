"""Developer micro-bench of the streaming operators beside the replay (measurement stream, comparison tracks,
raw-sensor pre-processing, RPY of a trajectory): achieved bytes/s against the measured copy bandwidth."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200.synth import make_imu

dev = torch.device("cuda:0")
N, T = 1 << 20, 100
base = make_imu(4096, T, seed=1, sigma=0.01, device=dev)
reps = N // 4096
streams = base.streams.repeat(1, 1, reps).contiguous()
acc_ref, mag_ref = base.acc_ref.repeat(1, reps).contiguous(), base.mag_ref.repeat(1, reps).contiguous()
tag = sys.argv[1] if len(sys.argv) > 1 else ""


def timed(fn, reps=4):
    best = 1e30
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i:
            best = min(best, e0.elapsed_time(e1))
    return best


def report(name, ms, nbytes):
    print(json.dumps({"tag": tag, "op": name, "ms": round(ms, 3), "gbs": round(nbytes / ms / 1e6, 1), "bytes_per_filter_step": nbytes // (N * T)}))


out = torch.empty_like(streams)
report("measurement_stream", timed(lambda: B.measurement_stream(streams, acc_ref, mag_ref, out=out)), N * T * (36 + 36))
report("tracks gyro+wahba", timed(lambda: B.tracks(streams, acc_ref, mag_ref, dt=0.01)), N * T * (36 + 32))
report("tracks wahba only", timed(lambda: B.tracks(streams, acc_ref, mag_ref, dt=0.01, want_gyro=False)), N * T * (24 + 16))
raw_prev = streams[:, 3:9].contiguous(); raw_next = (streams[:, 3:9] * 1.01).contiguous()
gyro = streams[:, 0:3].contiguous()
tspan = torch.rand((T, 4, N), device=dev) + 0.5
report("preprocess", timed(lambda: B.preprocess(gyro, raw_prev, raw_next, tspan, out=out)), N * T * (12 + 24 + 24 + 16 + 36))
traj = torch.randn((T, N, 4), device=dev)
traj /= traj.norm(dim=-1, keepdim=True)
report("traj2rpy", timed(lambda: B.traj2rpy(traj)), N * T * (16 + 12))
st = B.ReplayState.initial(N, dev, r=0.1)
tb = torch.empty((T, N, 4), device=dev)
report("replay + trajectory", timed(lambda: B.replay(streams, acc_ref, mag_ref, dt=0.01, q=1.0, r=0.1, state=B.ReplayState.initial(N, dev, r=0.1), out_traj=tb, precise_state=False)), N * T * (36 + 16))
