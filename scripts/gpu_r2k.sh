#!/bin/bash
# round 2, run k: thread-per-neuron kernels -- parity, sanitizer, PhysioNet sweeps
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family or thread_per_neuron or physionet or test_training_call or path_call" 2>&1 | tail -8
for w in physionet_synth_b50 physionet_synth_b2000 bs_demo_gru_5k; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2k_sweep_$w.json 2> gpurun_out/r2k_sweep_$w.err || tail -5 gpurun_out/r2k_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2k_sweep_$w.json
done
NJODE_NO_TPN=1 timeout 900 python bench.py --steps 5 --warmup 3 --workload physionet_synth_b50 --no-cpu-baseline --no-targets > gpurun_out/r2k_sweep_b50_notpn.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2k_sweep_b50_notpn.json
echo "=== memcheck tpn"
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn or thread_per_neuron" 2>&1 | grep -v "^=========     \|^  " | tail -6
echo "=== racecheck tpn"
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn and 50" 2>&1 | grep -v "^=========     \|^  " | tail -6
