#!/bin/bash
# r2b: whole-path units on the warp-GEMM kernels (njode_path.cuh): parity suite, sweeps of the whole-path workloads,
# eval path, ncu summaries of the new kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12
for w in physionet_synth_b50 physionet_synth_b2000 physionet_synth_b50_2x200 bs_demo_gru_5k; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2b_sweep_$w.json 2> gpurun_out/r2b_sweep_$w.err || tail -5 gpurun_out/r2b_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2b_sweep_$w.json
done
NJODE_NO_PATH=1 timeout 600 python bench.py --steps 3 --warmup 3 --workload physionet_synth_b50 --no-cpu-baseline --no-targets > gpurun_out/r2b_nopath_physionet_synth_b50.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2b_nopath_physionet_synth_b50.json
timeout 300 python scripts/eval_bench.py 4000 2>&1 | tail -6
for w in physionet_synth_b50 physionet_synth_b2000; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_path -c 2 -o gpurun_out/r2b_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2b_ncu_$w.err
  python scripts/ncu_summary.py gpurun_out/r2b_ncu_$w.ncu-rep > gpurun_out/r2b_ncu_$w.txt 2>&1
done
ls -la gpurun_out | tail -8
