#!/bin/bash
# round 2, run ad: glue warp prepares the next step's scalars and dropout layer keys (thread-per-neuron kernels, whole paths)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family or thread_per_neuron or physionet or path_call or gru" 2>&1 | tail -3
for w in physionet_synth_b50; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2ad_$w.json 2> gpurun_out/r2ad_$w.err || tail -5 gpurun_out/r2ad_$w.err
  python scripts/bench_line.py gpurun_out/r2ad_$w.json
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn and 50" 2>&1 | grep -v "^=========     \|^  " | tail -3
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn or thread_per_neuron" 2>&1 | grep -v "^=========     \|^  " | tail -3
