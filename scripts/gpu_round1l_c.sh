#!/bin/bash
# r1l (c): generic-kernel variants (row-loop layer GEMMs on/off) on the whole-path workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for v in 0 1; do
for w in physionet_synth_b50 physionet_synth_b2000 bs_2x100_5k; do
  NJODE_NO_ROWLOOP=$v timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r1l_c_${w}_$v.json 2> gpurun_out/r1l_c_${w}_$v.err || tail -5 gpurun_out/r1l_c_${w}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1l_c_${w}_$v.json")); r=d["roofline"]
    print("norowloop=$v %-24s B=%-6d S=%-5d ms/step %8.2f value %8.2fM e2e %8.2fM frac %.3f fwd %.2f bwd %.2f ms" % ("$w", d["config"]["paths_per_gpu"], d["config"]["euler_steps"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, r["frac"], r.get("fwd_kernel_ms", 0), r["kernel_ms"]))
except Exception as e: print("$w fail", e)
PY
done
done
