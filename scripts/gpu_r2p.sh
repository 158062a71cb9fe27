#!/bin/bash
# round 2, run p: segment thread-per-neuron kernels + shared gradient image -- parity, sanitizer, small-batch sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "segment or thread_per_neuron or every_kernel_family or recompute or demo_batch" 2>&1 | tail -5
for cfg in 0 1; do
  for w in bs_demo_200 bs_demo_1k bs_demo_5k; do
    NJODE_SEG_TPN=$cfg timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2p_${w}_tpn$cfg.json 2> gpurun_out/r2p_${w}_tpn$cfg.err || tail -5 gpurun_out/r2p_${w}_tpn$cfg.err
    python scripts/bench_line.py gpurun_out/r2p_${w}_tpn$cfg.json
  done
done
timeout 300 python bench.py --steps 5 --warmup 3 --workload physionet_synth_b50 --no-cpu-baseline --no-targets > gpurun_out/r2p_physionet_synth_b50.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2p_physionet_synth_b50.json
echo "=== racecheck + memcheck"
NJODE_SEG_TPN=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "segment_thread_per_neuron_kernels_train_mode and 200" 2>&1 | grep -v "^=========     \|^  " | tail -4
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "segment_thread_per_neuron or every_kernel_family and tpn" 2>&1 | grep -v "^=========     \|^  " | tail -4
