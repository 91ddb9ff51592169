"""Developer micro-bench of the Wahba-only entry point (BASELINE.json configs[3]: 100 M accel/mag pairs)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseestimationkf_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=100_000_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load()
M = a.m
g = torch.Generator(device=dev); g.manual_seed(1)
unit = lambda v: v / torch.linalg.vector_norm(v, dim=0, keepdim=True)
acc = unit(torch.randn((3, M), generator=g, device=dev)); mag = unit(torch.randn((3, M), generator=g, device=dev))
ra = torch.tensor([0.0, 0.0, 1.0], device=dev); rm = unit(torch.tensor([[0.4], [0.0], [-0.9165]], device=dev))[:, 0].contiguous()
qout = torch.empty((4, M), device=dev)
for wname, kw in (("half", (0.5, 0.5, 0)), ("reference", (0.0, 0.0, 1))):
    def run(n=M, out=qout):
        _lib.check(lib.posekf_wahba_f32(n, ra.data_ptr(), rm.data_ptr(), 1, acc.data_ptr(), mag.data_ptr(), None, None,
                                        kw[0], kw[1], kw[2], None, out.data_ptr(), _lib.WAHBA["qr2"], 0,
                                        torch.cuda.current_stream().cuda_stream), "wahba")
    best = 1e9
    for i in range(a.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        if i:
            best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"tag": a.tag, "weights": wname, "M": M, "ms": round(best, 4), "gsolves_per_s": round(M / best / 1e6, 2),
                      "hbm_gbs": round(M * 40 / best / 1e6, 1), "checksum": float(qout.double().abs().sum().item())}))
