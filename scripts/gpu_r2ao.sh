#!/bin/bash
# round 2, run ao: segment thread-per-neuron threshold between 400 and 1 000 paths
mkdir -p gpurun_out
for w in bs_demo_600; do for t in 0 1; do
  NJODE_SEG_TPN=$t timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2ao_${w}_tpn$t.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2ao_${w}_tpn$t.json
done; done
