#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -4
for w in bs_scaled_d16_h256_small bs_scaled_d16_h256; do
timeout 400 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -3 gpurun_out/bench_$w.err | grep -v -i warn | grep -v "return float"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json")); r=d["roofline"]
    print("$w", "ms/step %.2f value %.1fM e2e %.1fM frac %.3f | enc %.2f ode %.2f ro %.2f chain %.2f dw %.2f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, r["frac"], r["fwd_enc_ms"], r["fwd_ode_ms"], r["fwd_ro_ms"], r["bwd_chain_ms"], r["bwd_dw_ms"]))
except Exception as e: print("fail", e)
PY
done
if [ "$1" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nj_wide_kernel -s 3 -c 3 -o gpurun_out/prof_wide_fwd python bench.py --steps 1 --warmup 3 --workload bs_scaled_d16_h256_small --no-cpu-baseline > gpurun_out/ncu_wide.log 2>&1; tail -2 gpurun_out/ncu_wide.log
fi
