#!/bin/bash
# round 2, run ac: the library after the split into four njode_api*.cu translation units -- full GPU suite, smoke, three bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in heston_demo_20k bs_demo_200 physionet_synth_b50 physionet_synth_b300; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2ac_$w.json 2> gpurun_out/r2ac_$w.err || tail -5 gpurun_out/r2ac_$w.err
  python scripts/bench_line.py gpurun_out/r2ac_$w.json
done
