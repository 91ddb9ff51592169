#!/bin/bash
# On the GPU box: time every experiment build of the library in tools/variants/ (packed replay: libposekf_p*.so,
# Wahba-only: libposekf_w*.so); results under gpurun_out/
cd "$(dirname "$0")/.."
TAG=${1:-r01x}
mkdir -p gpurun_out
for so in tools/variants/libposekf_p*.so; do
  [ -e "$so" ] || continue
  tag=$(basename $so .so)
  POSEKF_LIB=$so python tools/devbench.py --t 500 --reps 5 --variants tma_packed:qr2,tma:qr2 --tag $tag 2>&1 | grep variant
done | tee gpurun_out/${TAG}_variants_packed.jsonl
for so in tools/variants/libposekf_w*.so; do
  [ -e "$so" ] || continue
  tag=$(basename $so .so)
  POSEKF_LIB=$so python tools/wahba_bench.py --tag $tag 2>&1 | grep weights
done | tee gpurun_out/${TAG}_variants_wahba.jsonl
