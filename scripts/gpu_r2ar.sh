#!/bin/bash
# round 2, run ar (4 GPUs): the driver's scaling launch at N = 4 on the last build
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2ar_bench_n4.json 2> gpurun_out/r2ar_bench_n4.err || tail -5 gpurun_out/r2ar_bench_n4.err
wc -l gpurun_out/r2ar_bench_n4.json; python scripts/bench_line.py gpurun_out/r2ar_bench_n4.json
python -c "
import json; d=json.load(open('gpurun_out/r2ar_bench_n4.json')); print(d['n_gpus'], d['value']/1e6, d['e2e']['value']/1e6, d.get('dp_check'))"
