#!/bin/bash
# runs tools/devbench.py against every library in tools/variants (on the GPU box)
cd "$(dirname "$0")/.."
V=${1:-tma:qr2,ldg:qr2}
for so in tools/variants/*.so; do
  tag=$(basename $so .so)
  POSEKF_LIB=$so python tools/devbench.py --t 500 --reps 4 --variants $V --tag $tag 2>&1 | grep variant
done
