#!/bin/bash
# round 2, run z2: the planner's choice between the kernel families at mid-size whole-path batches
mkdir -p gpurun_out
for w in physionet_synth_b300 physionet_synth_b600 bs_demo_gru_500; do
  for env in "A=1" "NJODE_NO_TPN=1" "NJODE_NO_TPN=1 NJODE_NO_STAT=1"; do
    env $env timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2z2_tmp.json 2> gpurun_out/r2z2_tmp.err || tail -5 gpurun_out/r2z2_tmp.err
    echo "[$env]"; python scripts/bench_line.py gpurun_out/r2z2_tmp.json
  done
done
