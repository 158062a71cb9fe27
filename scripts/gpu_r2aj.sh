#!/bin/bash
# round 2, run aj: the GPU parity suite under compute-sanitizer memcheck (full-size config-4 cases left out: thousands of steps)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_config4_physionet_full_size_against_oracle -k "not wide and not scaled and not train_script and not gpu_train" 2>&1 | grep -v "^=========     \|^  " | tail -8 | tee gpurun_out/r2aj_memcheck.txt
