#!/bin/bash
# usage: bash scripts/gpu_wide.sh case1 case2 ...
mkdir -p gpurun_out
for c in "$@"; do
  echo "=== $c"; timeout 70 python scripts/wide_debug.py $c 2>&1 | tail -25; echo "rc=$?"
done
