#!/bin/bash
# round 2, run al: class-A crossover (demo nets, GRU jump) between thread per neuron and the warp kernels
mkdir -p gpurun_out
for w in bs_demo_gru_1k bs_demo_gru_500; do for wv in 1 16; do
  NJODE_TPN_WAVES=$wv timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2al_${w}_waves$wv.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2al_${w}_waves$wv.json
done; done
