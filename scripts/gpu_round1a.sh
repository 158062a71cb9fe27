mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -3 gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a.json
python bench.py --steps 10 --warmup 3 --workload bs_demo_200 --no-cpu-baseline > gpurun_out/bench_r1a_bs200.json 2>> gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a_bs200.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nj_bwd_kernel -s 2 -c 1 -o gpurun_out/prof_bwd_r1a python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nj_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_fwd_r1a python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out
