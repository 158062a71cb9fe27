#!/bin/bash
# r2a: first GPU run of round 2 -- parity suite (new element-wise tolerance, full-size config 4, cond-exp fixtures),
# smoke, the default bench line with target_configs, reference arm, and the sweep lines the kernels work starts from
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench.err || tail -20 gpurun_out/r2a_bench.err
cut -c1-300 gpurun_out/r2a_bench_default.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2a_bench_reference_arm.json
for w in physionet_synth_b50 physionet_synth_b2000 bs_demo_gru_5k; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2a_sweep_$w.json 2> gpurun_out/r2a_sweep_$w.err || tail -5 gpurun_out/r2a_sweep_$w.err
  cut -c1-200 gpurun_out/r2a_sweep_$w.json
done
