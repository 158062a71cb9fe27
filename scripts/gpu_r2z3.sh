#!/bin/bash
# round 2, run z3: planner after the one-path-per-CTA rule; segment thread-per-neuron threshold at 400 paths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family or thread_per_neuron or physionet" 2>&1 | tail -3
for w in physionet_synth_b300 physionet_synth_b600 bs_demo_gru_500 physionet_synth_b50; do
  timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2z3_$w.json 2> gpurun_out/r2z3_$w.err || tail -5 gpurun_out/r2z3_$w.err
  python scripts/bench_line.py gpurun_out/r2z3_$w.json
done
for t in 0 1; do NJODE_SEG_TPN=$t timeout 600 python bench.py --steps 20 --warmup 5 --workload bs_demo_400 --no-cpu-baseline --no-targets > gpurun_out/r2z3_bs_demo_400_tpn$t.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2z3_bs_demo_400_tpn$t.json; done
