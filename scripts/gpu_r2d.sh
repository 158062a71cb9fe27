#!/bin/bash
mkdir -p gpurun_out
for r in 1 4; do NJODE_PATH_R=$r python scripts/dbg_path.py 12 40 2>&1 | tail -1 | cut -c1-400; done
echo "=== memcheck"
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python scripts/dbg_path.py 12 40 2>&1 | grep -v "^  " | cut -c1-300 | head -20
echo "=== tests"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8
for w in physionet_synth_b50 physionet_synth_b2000 bs_demo_gru_5k; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2d_sweep_$w.json 2> gpurun_out/r2d_sweep_$w.err || tail -5 gpurun_out/r2d_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2d_sweep_$w.json
done
for w in physionet_synth_b50 physionet_synth_b2000; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_path -c 2 -o gpurun_out/r2d_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2d_ncu_$w.err
  python scripts/ncu_summary.py gpurun_out/r2d_ncu_$w.ncu-rep > gpurun_out/r2d_ncu_$w.txt 2>&1
done
