#!/bin/bash
# round 2, run n: cooperative jump layers in the thread-per-neuron kernels -- parity (bounded by timeout: new barriers), sweep, sanitizer
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family or thread_per_neuron or physionet" 2>&1 | tail -3
for w in physionet_synth_b50; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2n_sweep_$w.json 2> gpurun_out/r2n_sweep_$w.err || tail -5 gpurun_out/r2n_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2n_sweep_$w.json
done
echo "=== racecheck + memcheck tpn"
timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn and 50" 2>&1 | grep -v "^=========     \|^  " | tail -4
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn or thread_per_neuron" 2>&1 | grep -v "^=========     \|^  " | tail -4
