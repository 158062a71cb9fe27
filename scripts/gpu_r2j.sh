#!/bin/bash
# round 2, run j: weight-stationary kernels for small segment batches -- parity, then the bs_demo_200 sweep (kernel families, tile heights)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -q -x -k "segment or recompute or demo_batch" 2>&1 | tail -8
for cfg in "0 0" "1 1" "1 2"; do
  set -- $cfg
  for w in bs_demo_200 bs_demo_1k; do
    NJODE_SEG_STAT=$1 NJODE_FORCE_TR=$2 timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2j_${w}_stat$1_tr$2.json 2> gpurun_out/r2j_${w}_stat$1_tr$2.err || tail -5 gpurun_out/r2j_${w}_stat$1_tr$2.err
    python scripts/bench_line.py gpurun_out/r2j_${w}_stat$1_tr$2.json
  done
done
for w in bs_demo_200 bs_demo_1k bs_demo_5k; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2j_${w}_auto.json 2> gpurun_out/r2j_${w}_auto.err || tail -5 gpurun_out/r2j_${w}_auto.err
  python scripts/bench_line.py gpurun_out/r2j_${w}_auto.json
done
echo "=== memcheck segstat"
NJODE_SEG_STAT=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "stationary_kernels_train_mode or both_kernel_families" 2>&1 | grep -v "^=========     \|^  " | tail -6
