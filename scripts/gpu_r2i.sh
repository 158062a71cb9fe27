#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
for w in bs_2x100_5k bs_2x100_20k bs_2x100_200; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2i_sweep_$w.json 2> gpurun_out/r2i_sweep_$w.err || tail -5 gpurun_out/r2i_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2i_sweep_$w.json
done
echo "=== memcheck: segment kernels with overflow tiles (2x100 nets), recompute mode"
NJODE_RECOMPUTE=on timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "2x100 or helper or demo_batch" 2>&1 | grep -v "^=========     \|^  " | tail -6
