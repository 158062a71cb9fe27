#!/bin/bash
# round 2, final evidence pass on the last build: full GPU suite, smoke, every workload, default bench + reference arm, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in heston_demo_20k bs_demo_200 bs_demo_400 bs_demo_1k bs_demo_5k ou_demo_20k hestonwof_demo_1k hestonwof_demo_20k bs_2x100_200 bs_2x100_5k bs_2x100_20k bs_demo_gru_500 bs_demo_gru_5k physionet_synth_b50 physionet_synth_b300 physionet_synth_b600 physionet_synth_b2000 physionet_synth_b50_2x200; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-targets > gpurun_out/r2final_sweep_$w.json 2> gpurun_out/r2final_sweep_$w.err || tail -5 gpurun_out/r2final_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2final_sweep_$w.json
done
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2final_reference_arm.json 2>/dev/null; wc -l gpurun_out/r2final_reference_arm.json
timeout 900 python bench.py > gpurun_out/r2final_bench_default.json 2> gpurun_out/r2final_bench_default.err; wc -l gpurun_out/r2final_bench_default.json; python scripts/bench_line.py gpurun_out/r2final_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-targets > /dev/null 2>&1
