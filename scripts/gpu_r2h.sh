#!/bin/bash
# r2h: full GPU suite, smoke, default bench line (with target_configs + recompute record), reference arm, sweeps
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench.err || tail -20 gpurun_out/r2h_bench.err
python scripts/bench_line.py gpurun_out/r2h_bench_default.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_default.json"))
print("recompute:", d.get("recompute"))
for k,t in d.get("target_configs",{}).items():
    print(k, "ms/step %.3f value %.2fM e2e %.2fM frac %.4f" % (t["ms_per_step"], t["value"]/1e6, t["e2e"]["value"]/1e6, t["roofline"]["frac"]), "speedups", t.get("speedup_e2e_vs_cpu_all_threads"), t.get("speedup_e2e_vs_cpu_1thread"), "gen", t.get("generator"))
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2h_bench_reference_arm.json
for w in bs_demo_200 ou_demo_20k hestonwof_demo_1k bs_2x100_5k bs_2x100_20k physionet_synth_b50 physionet_synth_b2000 physionet_synth_b50_2x200 bs_demo_gru_5k; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2h_sweep_$w.json 2> gpurun_out/r2h_sweep_$w.err || tail -5 gpurun_out/r2h_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2h_sweep_$w.json
done
