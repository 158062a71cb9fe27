#!/bin/bash
# round 2, run z: thread-per-neuron kernels with several tiles per CTA (forced) against the pipelined warp kernels at larger whole-path batches
mkdir -p gpurun_out
for t in 0 1; do for w in physionet_synth_b2000 bs_demo_gru_5k; do
  NJODE_FORCE_TPN=$t timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2z_${w}_tpn$t.json 2> gpurun_out/r2z_${w}_tpn$t.err || tail -5 gpurun_out/r2z_${w}_tpn$t.err
  python scripts/bench_line.py gpurun_out/r2z_${w}_tpn$t.json
done; done
