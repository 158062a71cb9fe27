#!/bin/bash
# round 2, run x: saved hidden activations of the ODE network (segment kernels) on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "segment or demo_batch or training_call or recompute or 2x100 or helper or config3" 2>&1 | tail -3
for sv in 1 0; do for w in heston_demo_20k bs_demo_5k bs_2x100_5k bs_demo_1k; do
  NJODE_SAVE_ACTIVATIONS_MAX_MB=4096 NJODE_SAVE_ACTIVATIONS=$sv timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2x_${w}_save$sv.json 2> gpurun_out/r2x_${w}_save$sv.err || tail -5 gpurun_out/r2x_${w}_save$sv.err
  python scripts/bench_line.py gpurun_out/r2x_${w}_save$sv.json
done; done
