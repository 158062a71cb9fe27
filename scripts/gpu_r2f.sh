#!/bin/bash
# r2f: evidence runs -- f1 test detail, mma.sync probe, training curves, tensor-core memcheck + tensor-pipe counter
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sde.py -q -x -k "no_host_round_trip" --tb=short 2>&1 | tail -25
python scripts/mma_peak.py
timeout 900 python scripts/train_demo.py --epochs 24 --out gpurun_out/r2_train_bs_demo.json 2>&1 | tail -26
timeout 900 python scripts/train_demo.py --compare-tensor-cores --epochs 4 --out gpurun_out/r2_train_tensor_cores.json 2>&1 | tail -12
echo "=== memcheck over the tensor-core tests"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_wide.py -q -x 2>&1 | grep -v "^=========     \|^  " | tail -12
echo "=== ncu tensor pipe, small scaled workload"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:nj_wide -c 40 --csv --log-file gpurun_out/r2f_ncu_wide_tensor.csv python bench.py --steps 1 --warmup 0 --workload bs_scaled_d16_h256_small --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2f_ncu_wide.err
tail -3 gpurun_out/r2f_ncu_wide.err; wc -l gpurun_out/r2f_ncu_wide_tensor.csv
