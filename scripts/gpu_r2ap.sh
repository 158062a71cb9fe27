#!/bin/bash
# round 2, run ap: return_path (evaluation) calls at small batch sizes: thread per neuron with per-step records vs the warp kernels
mkdir -p gpurun_out
for B in 100 500; do for t in 0 1; do
  if [ $t = 0 ]; then e="NJODE_NO_TPN=1 NJODE_NO_STAT=1"; else e="A=1"; fi
  echo "== B=$B [$e]"; env $e timeout 300 python scripts/eval_bench.py $B 2>/dev/null | grep "outputs stay\|evaluate"
done; done
