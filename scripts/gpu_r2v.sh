#!/bin/bash
# round 2, run v: final state -- full GPU suite, smoke, every sweep workload, launch list + ncu captures of the default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in heston_demo_20k bs_demo_200 bs_demo_1k bs_demo_5k ou_demo_20k hestonwof_demo_1k hestonwof_demo_20k bs_2x100_200 bs_2x100_5k bs_2x100_20k bs_demo_gru_5k physionet_synth_b50 physionet_synth_b50_2x200 physionet_synth_b2000; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-targets > gpurun_out/r2v_sweep_$w.json 2> gpurun_out/r2v_sweep_$w.err || tail -5 gpurun_out/r2v_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2v_sweep_$w.json
done
timeout 900 python scripts/eval_bench.py > gpurun_out/r2v_eval.txt 2>&1; tail -4 gpurun_out/r2v_eval.txt
# launch list of the default bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-targets > /dev/null 2>&1
for w in heston_demo_20k; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_seg -c 2 -o gpurun_out/r2v_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2v_ncu_$w.err
  python scripts/ncu_summary.py gpurun_out/r2v_ncu_$w.ncu-rep > gpurun_out/r2v_ncu_$w.txt 2>&1
done
