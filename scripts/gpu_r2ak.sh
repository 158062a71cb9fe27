#!/bin/bash
# round 2, run ak: racecheck over the parity tests of every kernel family (small cases), memcheck of the wide + full-size cases
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "test_training_call_with_hT_gradient or test_path_call or every_kernel_family or segment_thread_per_neuron or gru_jump or helper or recompute_mode_golden" 2>&1 | grep -v "^=========     \|^  " | tail -6 | tee gpurun_out/r2ak_racecheck.txt
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -q -x -k "wide or config4 or config5" 2>&1 | grep -v "^=========     \|^  " | tail -6 | tee gpurun_out/r2ak_memcheck_wide_fullsize.txt
