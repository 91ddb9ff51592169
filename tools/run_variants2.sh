#!/bin/bash
# On the GPU box: packed-replay tile/occupancy variants and Wahba-only variants (tools/variants/*.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for so in tools/variants/libposekf_p*.so; do
  tag=$(basename $so .so)
  POSEKF_LIB=$so python tools/devbench.py --t 500 --reps 5 --variants tma_packed:qr2 --tag $tag 2>&1 | grep variant
done | tee gpurun_out/r01h_variants_packed.jsonl
for so in tools/variants/libposekf_w*.so; do
  tag=$(basename $so .so)
  POSEKF_LIB=$so python tools/wahba_bench.py --tag $tag 2>&1 | grep weights
done | tee gpurun_out/r01h_variants_wahba.jsonl
