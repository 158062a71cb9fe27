#!/bin/bash
# round 2, run y (8 GPUs): the driver's scaling launch at N = 8, both arms
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err || tail -5 gpurun_out/r2y_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2y_ref_n8.json 2> gpurun_out/r2y_ref_n8.err; wc -l gpurun_out/r2y_ref_n8.json
python - <<'PY'
import json
f="gpurun_out/r2y_bench_n8.json"
txt=open(f).read().strip().splitlines()
print("stdout lines:", len(txt))
d=json.loads(txt[-1])
print("n_gpus", d["n_gpus"], "value %.1fM e2e %.1fM ms %.2f launches %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["gpu_launches"]), "dp_check", d.get("dp_check"))
for k,v in (d.get("target_configs") or {}).items(): print("   target", k, v.get("ms_per_step"), v.get("value"), v.get("e2e",{}).get("value"))
PY
