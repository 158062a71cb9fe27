#!/bin/bash
# usage: bash scripts/gpu_check.sh TAG [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
python bench.py --steps 10 --warmup 3 --workload bs_demo_200 --no-cpu-baseline > gpurun_out/bench_${TAG}_bs200.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_${TAG}_bs200.json
if [ "$2" == "ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_kernel -s 2 -c 1 -o gpurun_out/prof_bwd_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 2 -c 1 -o gpurun_out/prof_fwd_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
fi
