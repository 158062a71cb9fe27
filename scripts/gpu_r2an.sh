#!/bin/bash
# round 2, run an: ncu summary (no source import: small report) of the last build's thread-per-neuron kernels at 300 PhysioNet records
mkdir -p gpurun_out
w=physionet_synth_b300
timeout 900 ncu --set full --clock-control none -k regex:nj_tpn -c 2 -o gpurun_out/r2an_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2an_ncu_$w.err
python scripts/ncu_summary.py gpurun_out/r2an_ncu_$w.ncu-rep > gpurun_out/r2an_ncu_$w.txt 2>&1
ls -la gpurun_out | grep r2an
