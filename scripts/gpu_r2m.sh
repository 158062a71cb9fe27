#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family or thread_per_neuron or physionet" 2>&1 | tail -3
for w in physionet_synth_b50; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2m_sweep_$w.json 2> gpurun_out/r2m_sweep_$w.err || tail -5 gpurun_out/r2m_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2m_sweep_$w.json
done
