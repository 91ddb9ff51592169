"""Small end-to-end workload for compute-sanitizer (memcheck): every replay variant on ragged shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseestimationkf_b200 import batched as B
from poseestimationkf_b200.synth import make_imu
dev = torch.device("cuda:0")
for N, T in ((1000, 37), (130, 9), (516, 5)):
    imu = make_imu(N, T, seed=N, sigma=0.01, device=dev, keep_truth=True)
    truth = imu.q_true.permute(0, 2, 1).to(torch.float32).contiguous()
    for staging in (("tma", "ldg") if N % 4 == 0 else ("auto", "ldg")):
        for wahba in ("qr2", "jacobi"):
            for precise in (False, True):
                st, traj, fl = B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, store_trajectory=True, store_flips=True,
                                        truth=truth, lpf_alpha_acc=0.1, lpf_alpha_mag=0.1, precise_state=precise, wahba=wahba,
                                        staging=staging)
                B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, precise_state=precise, wahba=wahba, staging=staging)
    B.tracks(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt)
    B.wahba(imu.acc_ref, imu.mag_ref, imu.streams[0, 3:6].contiguous(), imu.streams[0, 6:9].contiguous(), weights_from_acc=True, algo="jacobi")
sw = make_imu(256, 11, seed=1, device=dev)
B.replay(sw.streams, sw.acc_ref, sw.mag_ref, dt=0.01, q=torch.ones(1024, device=dev), r=torch.full((1024,), 0.1, device=dev), n_filters=1024)
host = sw.streams.cpu().pin_memory()
B.replay_host(host, sw.acc_ref.cpu(), sw.mag_ref.cpu(), dt=0.01, q=torch.ones(256), r=torch.full((256,), 0.1), store_trajectory=True, chunk_steps=4)
torch.cuda.synchronize()
print("sanitize target finished OK")
