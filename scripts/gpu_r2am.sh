#!/bin/bash
# round 2, run am: tile-height classes of the segment kernels on the default bench workload (forced heights against the planner's mix)
mkdir -p gpurun_out
for tr in 0 1 2 4; do
  if [ $tr = 0 ]; then e="A=1"; else e="NJODE_FORCE_TR=$tr"; fi
  env $e timeout 600 python bench.py --steps 10 --warmup 3 --workload heston_demo_20k --no-cpu-baseline --no-targets > gpurun_out/r2am_tr$tr.json 2>/dev/null; echo "[$e]"; python scripts/bench_line.py gpurun_out/r2am_tr$tr.json
done
for nw in 8 10; do
  NJODE_FORCE_NW=$nw timeout 600 python bench.py --steps 10 --warmup 3 --workload heston_demo_20k --no-cpu-baseline --no-targets > gpurun_out/r2am_nw$nw.json 2>/dev/null; echo "[NW=$nw]"; python scripts/bench_line.py gpurun_out/r2am_nw$nw.json
done
