#!/bin/bash
# round 2, run ae (2 GPUs): last build -- full GPU suite, smoke, default bench, and the N = 2 launch (stdout must be ONE line)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2ae_bench_default.json 2> gpurun_out/r2ae_bench_default.err; wc -l gpurun_out/r2ae_bench_default.json; python scripts/bench_line.py gpurun_out/r2ae_bench_default.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ae_bench_n2.json 2> gpurun_out/r2ae_bench_n2.err || tail -5 gpurun_out/r2ae_bench_n2.err
wc -l gpurun_out/r2ae_bench_n2.json; python scripts/bench_line.py gpurun_out/r2ae_bench_n2.json
for w in physionet_synth_b50 bs_demo_200; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-targets > gpurun_out/r2ae_sweep_$w.json 2>/dev/null; python scripts/bench_line.py gpurun_out/r2ae_sweep_$w.json
done
