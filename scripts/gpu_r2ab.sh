#!/bin/bash
# round 2, run ab: source-level ncu capture of the jump-heavy whole-path kernels (current build)
mkdir -p gpurun_out
w=physionet_synth_b300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_path -c 2 -o gpurun_out/r2ab_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2ab_ncu_$w.err
python scripts/ncu_summary.py gpurun_out/r2ab_ncu_$w.ncu-rep > gpurun_out/r2ab_ncu_$w.txt 2>&1
ls -la gpurun_out
