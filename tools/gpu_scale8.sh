#!/bin/bash
# 8-GPU visit: the contract bench line at N=8 (NUMA-bound ranks) and a short run without the binding for the e2e comparison
cd "$(dirname "$0")/.."
TAG=${1:-r01x}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8"
timeout 600 $RUN --steps 10 --warmup 3 --no-variants --cpu-seconds 5 > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_bench_8gpu.err; echo "bound exit $?"
timeout 400 $RUN --steps 3 --warmup 3 --timesteps 250 --no-variants --no-cpu-baseline --no-numa-bind > gpurun_out/${TAG}_bench_8gpu_nobind.json 2> gpurun_out/${TAG}_bench_8gpu_nobind.err; echo "nobind exit $?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_8gpu.json", "gpurun_out/${TAG}_bench_8gpu_nobind.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"]/1e9, "G; e2e", d["e2e"]["value"]/1e9, "G", d["e2e"].get("host_affinity"), d["e2e"]["h2d_gbs"])
    except Exception as e:
        print(f, "unreadable", e)
PY
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
