#!/bin/bash
# round 2, run r: after the FFMA-chain / unrolled-loss-loop changes -- full GPU suite + the small-batch sweeps
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
for w in physionet_synth_b50 bs_demo_200 physionet_synth_b2000 bs_demo_gru_5k; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2r_sweep_$w.json 2> gpurun_out/r2r_sweep_$w.err || tail -5 gpurun_out/r2r_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2r_sweep_$w.json
done
