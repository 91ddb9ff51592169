"""Two-GPU sharded replay + NCCL gather of the final states equals the single-GPU replay bit for bit
(filters are independent; shard boundaries are 128-aligned).  Skipped unless >= 2 GPUs are visible."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, t, out_dir):
    from poseestimationkf_b200 import batched as B
    from poseestimationkf_b200 import sharding as SH
    from poseestimationkf_b200.synth import make_imu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        imu = make_imu(n, t, seed=123, sigma=0.01, device=dev)          # same seed -> same data on every rank
        b, e = SH.shard_bounds(n, rank, world)
        st, _, _ = B.replay(imu.streams[:, :, b:e].contiguous(), imu.acc_ref[:, b:e].contiguous(),
                            imu.mag_ref[:, b:e].contiguous(), dt=imu.dt, q=1.0, r=0.1)
        full_x = SH.gather_states(st.x, n)
        full_p = SH.gather_states(st.p, n)
        slowest = SH.max_over_ranks(float(rank), dev)
        if rank == 0:
            ref, _, _ = B.replay(imu.streams, imu.acc_ref, imu.mag_ref, dt=imu.dt, q=1.0, r=0.1)
            torch.save({"x": torch.equal(full_x, ref.x), "p": torch.equal(full_p, ref.p), "slowest": slowest,
                        "sizes": SH.shard_sizes(n, world)}, os.path.join(out_dir, "res.pt"))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_replay_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, t = 1000 * 128 + 72, 40            # ragged: the last shard is shorter and not a multiple of 128
    mp.spawn(_worker, args=(2, _free_port(), n, t, str(tmp_path)), nprocs=2, join=True)
    res = torch.load(os.path.join(str(tmp_path), "res.pt"))
    assert res["x"] and res["p"] and res["slowest"] == 1.0 and sum(res["sizes"]) == n
