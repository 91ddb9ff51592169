#!/bin/bash
# N-GPU visit (N = 2 or 4): the contract bench line at N GPUs
cd "$(dirname "$0")/.."
N=$1; TAG=${2:-r01x}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-variants --cpu-seconds 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value']/1e9, 'G; e2e', d['e2e']['value']/1e9)"
