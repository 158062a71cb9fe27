#!/bin/bash
# round 2, run s (2 GPUs): the driver's launch of bench.py at N = 2 (torchrun), both arms, and the default N = 1 line
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2s_ref_n1.json 2> gpurun_out/r2s_ref_n1.err; tail -c 600 gpurun_out/r2s_ref_n1.json; echo
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err || tail -5 gpurun_out/r2s_bench_n1.err
python scripts/bench_line.py gpurun_out/r2s_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err || tail -5 gpurun_out/r2s_bench_n2.err
python scripts/bench_line.py gpurun_out/r2s_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2s_ref_n2.json 2> gpurun_out/r2s_ref_n2.err; tail -c 300 gpurun_out/r2s_ref_n2.json; echo
python - <<'PY'
import json
for f in ("gpurun_out/r2s_bench_n1.json","gpurun_out/r2s_bench_n2.json"):
    try:
        d=json.load(open(f)); print(f, "n_gpus", d["n_gpus"], "value %.1fM e2e %.1fM ms %.2f launches %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["gpu_launches"]), "dp_check", d.get("dp_check"))
        for k,v in (d.get("target_configs") or {}).items(): print("   target", k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("ms_per_step","value","speedup_e2e_vs_cpu_all_threads","speedup_e2e_vs_cpu_1thread")}, "e2e", v.get("e2e",{}).get("value"))
    except Exception as e: print(f, "FAIL", e)
PY
