#!/bin/bash
# One GPU-box visit that refreshes the evidence under gpurun_out/: GPU tests, smoke, parity report, the default bench
# line, the ncu launch list of the same (shortened) command, and one `ncu --set full` capture of the packed replay kernel.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh TAG'
cd "$(dirname "$0")/.."
TAG=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
python tests/parity_report.py gpurun_out/${TAG}_parity_report.json > /dev/null 2> gpurun_out/${TAG}_parity_report.err; echo "parity report exit $?"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json'))
print(d['value']/1e9, 'G', d['roofline']['frac'], d['e2e']['value']/1e9, d['clocks'], d.get('variants_gsteps_per_s_1gpu_250_timesteps'))"
SHORT="python bench.py --steps 3 --warmup 3 --timesteps 200 --no-e2e --no-cpu-baseline --no-variants"
$SHORT > gpurun_out/${TAG}_bench_short.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/${TAG}_launch_list.csv $SHORT > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:replay_tma2 -c 1 -f \
    -o gpurun_out/${TAG}_replay_packed $SHORT > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -12
