#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 4 8; do NJODE_PATH_R=$r python scripts/dbg_path.py 12 40 2>&1 | tail -1 | cut -c1-330; done
echo "=== memcheck"; timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python scripts/dbg_path.py 12 40 2>&1 | grep -v "^  " | cut -c1-200 | head -12
echo "=== racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python scripts/dbg_path.py 6 20 2>&1 | grep -v "^  " | cut -c1-200 | head -12
echo "=== tests"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
for w in physionet_synth_b50 physionet_synth_b2000 bs_demo_gru_5k; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2e_sweep_$w.json 2> gpurun_out/r2e_sweep_$w.err || tail -5 gpurun_out/r2e_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2e_sweep_$w.json
done
timeout 300 python scripts/eval_bench.py 4000 2>&1 | tail -5
w=physionet_synth_b50
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_stat -c 2 -o gpurun_out/r2e_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2e_ncu_$w.err
python scripts/ncu_summary.py gpurun_out/r2e_ncu_$w.ncu-rep > gpurun_out/r2e_ncu_$w.txt 2>&1
NJODE_FORCE_STAT=1 timeout 600 python bench.py --steps 3 --warmup 3 --workload physionet_synth_b2000 --no-cpu-baseline --no-targets > gpurun_out/r2e_stat_physionet_synth_b2000.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2e_stat_physionet_synth_b2000.json
