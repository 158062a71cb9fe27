#!/bin/bash
# round 2, run aa: dW phases of a jump walk only their network's tile range
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family or thread_per_neuron or physionet or 2x100 or helper or gru" 2>&1 | tail -3
for w in physionet_synth_b50 physionet_synth_b300 physionet_synth_b2000 bs_demo_gru_5k bs_2x100_5k bs_demo_200 heston_demo_20k; do
  timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2aa_$w.json 2> gpurun_out/r2aa_$w.err || tail -5 gpurun_out/r2aa_$w.err
  python scripts/bench_line.py gpurun_out/r2aa_$w.json
done
