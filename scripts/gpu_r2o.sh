#!/bin/bash
# round 2, run o: ncu capture (full set + source) of the thread-per-neuron kernels on the PhysioNet batch of 50
mkdir -p gpurun_out
w=physionet_synth_b50
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_tpn -c 2 -o gpurun_out/r2o_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2o_ncu_$w.err
python scripts/ncu_summary.py gpurun_out/r2o_ncu_$w.ncu-rep > gpurun_out/r2o_ncu_$w.txt 2>&1
tail -30 gpurun_out/r2o_ncu_$w.txt
