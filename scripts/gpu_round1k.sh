#!/bin/bash
# full GPU evidence run: tests, default bench, config-5 bench, ncu launch lists + full captures of the wide kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -2 gpurun_out/bench_r1k.err | grep -v -i warn; cut -c1-400 gpurun_out/bench_r1k.json
timeout 600 python bench.py --steps 5 --warmup 3 --workload bs_scaled_d16_h256 > gpurun_out/bench_r1k_cfg5.json 2> gpurun_out/bench_r1k_cfg5.err; tail -2 gpurun_out/bench_r1k_cfg5.err | grep -v -i warn; cut -c1-300 gpurun_out/bench_r1k_cfg5.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1k_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1k_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/launches_r1k_cfg5small.csv python bench.py --steps 2 --warmup 3 --workload bs_scaled_d16_h256_small --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_wide -s 9 -c 9 -o gpurun_out/prof_wide_r1k python bench.py --steps 1 --warmup 3 --workload bs_scaled_d16_h256_small --no-cpu-baseline > gpurun_out/ncu_w.log 2>&1; tail -2 gpurun_out/ncu_w.log
