#!/bin/bash
mkdir -p gpurun_out
NJODE_NO_STAT=1 python scripts/dbg_path.py 12 40 2>&1 | tail -1 | cut -c1-330
echo "=== racecheck (pipelined backward)"; NJODE_NO_STAT=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python scripts/dbg_path.py 6 20 2>&1 | grep -v "^  " | cut -c1-200 | head -8
echo "=== tests"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5
for w in physionet_synth_b2000 bs_demo_gru_5k physionet_synth_b50; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2g_sweep_$w.json 2> gpurun_out/r2g_sweep_$w.err || tail -5 gpurun_out/r2g_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2g_sweep_$w.json
done
for r in 4 8; do
NJODE_PATH_R=$r timeout 600 python bench.py --steps 3 --warmup 3 --workload physionet_synth_b2000 --no-cpu-baseline --no-targets > gpurun_out/r2g_R${r}_physionet_synth_b2000.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2g_R${r}_physionet_synth_b2000.json
done
NJODE_NO_PIPE=1 timeout 600 python bench.py --steps 3 --warmup 3 --workload bs_demo_gru_5k --no-cpu-baseline --no-targets > gpurun_out/r2g_nopipe_gru.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2g_nopipe_gru.json
w=physionet_synth_b2000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_path -c 2 -o gpurun_out/r2g_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2g_ncu_$w.err
python scripts/ncu_summary.py gpurun_out/r2g_ncu_$w.ncu-rep > gpurun_out/r2g_ncu_$w.txt 2>&1
