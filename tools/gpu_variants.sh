#!/bin/bash
# On the GPU box: time every experiment build in tools/variants/ with tools/devbench.py (packed kernel, 1 Mi x 500).
# usage: tools/gpu_variants.sh TAG [glob]      results -> gpurun_out/TAG_variants.jsonl
cd "$(dirname "$0")/.."
TAG=${1:-r02x}
GLOB=${2:-libposekf_*.so}
mkdir -p gpurun_out
for so in tools/variants/$GLOB; do
  [ -e "$so" ] || continue
  tag=$(basename $so .so)
  POSEKF_LIB=$so timeout 300 python tools/devbench.py --t 500 --reps 5 --sustain ${SUSTAIN:-0} --variants tma_packed:qr2 --tag $tag 2>&1 | grep variant
done | tee gpurun_out/${TAG}_variants.jsonl
