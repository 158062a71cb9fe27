#!/bin/bash
# usage: bash scripts/gpu_wide.sh case1 case2 ...   (full logs in gpurun_out/wide_<case>.log)
mkdir -p gpurun_out
for c in "$@"; do
  echo "=== $c"; timeout 150 python tests/tools/wide_debug.py $c > gpurun_out/wide_$c.log 2>&1; echo "rc=$?"
  grep -v -E "^frame|^$" gpurun_out/wide_$c.log | tail -45
done
