#!/bin/bash
# round 2, run t: segment backward without the overflow code where every dW tile has a register slot
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "segment or demo_batch or training_call or recompute or 2x100 or helper" 2>&1 | tail -3
for w in heston_demo_20k bs_demo_5k bs_2x100_5k; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2t_sweep_$w.json 2> gpurun_out/r2t_sweep_$w.err || tail -5 gpurun_out/r2t_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2t_sweep_$w.json
done
