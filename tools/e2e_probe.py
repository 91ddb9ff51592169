"""Host-link probe: plain pinned H2D rate vs posekf_replay_host_f32 at several sizes / chunkings."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200.synth import make_imu
dev = torch.device("cuda:0")
N = 1 << 20
base = make_imu(4096, 500, seed=1, sigma=0.01, device=dev)
reps = N // 4096
acc_ref = base.acc_ref.repeat(1, reps).cpu().pin_memory(); mag_ref = base.mag_ref.repeat(1, reps).cpu().pin_memory()
q = torch.full((N,), 1.0).pin_memory(); r = torch.full((N,), 0.1).pin_memory()
for T in [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else (125, 250, 500):
    host = torch.empty((T, 9, N), dtype=torch.float32, pin_memory=True)
    host.copy_(base.streams[:T].repeat(1, 1, reps))
    scratch = torch.empty((T, 9, N), dtype=torch.float32, device=dev)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); scratch.copy_(host, non_blocking=True); torch.cuda.synchronize()
        plain = host.numel() * 4 / (time.perf_counter() - t0) / 1e9
    del scratch
    for chunk in [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else (0, 4, 30, -1):
        ws = B.HostWorkspace(N, chunk_steps=chunk) if chunk >= 0 else None
        B.replay_host(host, acc_ref, mag_ref, dt=0.01, q=q, r=r, workspace=ws)
        dts = []
        for _ in range(3):
            t0 = time.perf_counter()
            B.replay_host(host, acc_ref, mag_ref, dt=0.01, q=q, r=r, workspace=ws)
            dts.append(time.perf_counter() - t0)
        dt = min(dts)
        if ws: ws.close()
        print(json.dumps({"T": T, "chunk_steps": chunk, "plain_h2d_gbs": round(plain, 1), "replay_host_gbs": round(host.numel() * 4 / dt / 1e9, 1),
                          "seconds": round(dt, 4), "all_seconds": [round(d, 4) for d in dts], "gsteps_per_s": round(N * T / dt / 1e9, 3)}))
    del host
