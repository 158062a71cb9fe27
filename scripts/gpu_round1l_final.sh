#!/bin/bash
# r1l final evidence run (no profiler): GPU parity suite, smoke, bench lines of every workload with the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r1l_bench_heston_demo_20k.json 2> gpurun_out/r1l_bench.err; cut -c1-200 gpurun_out/r1l_bench_heston_demo_20k.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1l_bench_reference_arm.json 2>/dev/null; cut -c1-160 gpurun_out/r1l_bench_reference_arm.json
timeout 600 python bench.py --steps 5 --warmup 3 --workload bs_scaled_d16_h256 --no-cpu-baseline > gpurun_out/r1l_bench_bs_scaled_d16_h256.json 2> gpurun_out/r1l_bench_cfg5.err; cut -c1-200 gpurun_out/r1l_bench_bs_scaled_d16_h256.json
for w in bs_demo_200 ou_demo_20k hestonwof_demo_1k hestonwof_demo_20k bs_2x100_5k bs_demo_gru_5k physionet_synth_b50 physionet_synth_b2000; do
  extra="--no-cpu-baseline"
  if [ "$w" == "bs_demo_200" ]; then extra=""; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w $extra > gpurun_out/r1l_sweep_$w.json 2> gpurun_out/r1l_sweep_$w.err || tail -5 gpurun_out/r1l_sweep_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1l_sweep_$w.json")); r=d["roofline"]; cb=d.get("cpu_baseline")
    print("%-24s B=%-6d S=%-5d ms/step %8.2f value %8.2fM e2e %8.2fM frac %.3f fwd %.2f bwd %.2f ms cpu %s" % ("$w", d["config"]["paths_per_gpu"], d["config"]["euler_steps"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, r["frac"], r.get("fwd_kernel_ms", 0), r["kernel_ms"], ("%.1fk" % (cb["value"]/1e3)) if cb else "-"))
except Exception as e: print("$w fail", e)
PY
done
