#!/bin/bash
# round 2, run af: thread-per-neuron kernels with several one-path tiles per CTA against the pipelined warp kernels
mkdir -p gpurun_out
for w in physionet_synth_b300 physionet_synth_b600 physionet_synth_b2000 bs_demo_gru_500 bs_demo_gru_5k; do
  for wv in 1 16; do
    NJODE_TPN_WAVES=$wv timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2af_${w}_waves$wv.json 2> gpurun_out/r2af_${w}_waves$wv.err || tail -5 gpurun_out/r2af_${w}_waves$wv.err
    python scripts/bench_line.py gpurun_out/r2af_${w}_waves$wv.json
  done
done
