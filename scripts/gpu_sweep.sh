#!/bin/bash
# measured table of the parity-test configurations (BASELINE configs 1-4): one bench line per workload
mkdir -p gpurun_out
for w in bs_demo_200 ou_demo_20k hestonwof_demo_1k hestonwof_demo_20k bs_2x100_5k physionet_synth_b50 physionet_synth_b2000; do
  extra="--no-cpu-baseline"
  if [ "$w" == "physionet_synth_b50" ] || [ "$w" == "bs_demo_200" ]; then extra=""; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w $extra > gpurun_out/sweep_$w.json 2> gpurun_out/sweep_$w.err || tail -5 gpurun_out/sweep_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$w.json")); r=d["roofline"]; cb=d.get("cpu_baseline")
    print("%-24s B=%-6d S=%-5d ms/step %8.2f value %8.2fM e2e %8.2fM frac %.3f fwd %.2f bwd %.2f ms cpu %s" % ("$w", d["config"]["paths_per_gpu"], d["config"]["euler_steps"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, r["frac"], r.get("fwd_kernel_ms", 0), r["kernel_ms"], ("%.1fk" % (cb["value"]/1e3)) if cb else "-"))
except Exception as e: print("$w fail", e)
PY
done
