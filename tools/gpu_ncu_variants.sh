#!/bin/bash
# ncu --set full captures of the other two packed replay kernels (trajectory output, precise variant)
cd "$(dirname "$0")/.."
TAG=${1:-r01x}
mkdir -p gpurun_out
python tools/devbench.py --t 200 --reps 2 --variants tma_packed:qr2 --traj 1 > gpurun_out/${TAG}_devbench_traj.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:replay_tma2 -c 1 -f -o gpurun_out/${TAG}_replay_packed_traj \
    python tools/devbench.py --t 200 --reps 1 --variants tma_packed:qr2 --traj 1 > gpurun_out/${TAG}_ncu_traj.log 2>&1; echo "ncu traj exit $?"
SHORT="python bench.py --steps 1 --warmup 3 --timesteps 200 --no-e2e --no-cpu-baseline --no-variants --precise-state"
$SHORT > gpurun_out/${TAG}_bench_precise_short.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:replay_tma2 -c 1 -f \
    -o gpurun_out/${TAG}_replay_packed_precise $SHORT > gpurun_out/${TAG}_ncu_precise.log 2>&1; echo "ncu precise exit $?"
ls -la gpurun_out | grep ${TAG}
