#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for t in 0 1; do
NJODE_SEG_TPN=$t timeout 300 python bench.py --steps 30 --warmup 5 --workload bs_demo_200 --no-cpu-baseline --no-targets > gpurun_out/r2q_t${t}_$rep.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2q_t${t}_$rep.json
done; done
