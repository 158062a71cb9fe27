#!/bin/bash
# round 2, run u: cooperative layers in the segment thread-per-neuron kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "segment or demo_batch or recompute" 2>&1 | tail -3
for w in bs_demo_200; do
  timeout 300 python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2u_sweep_$w.json 2> gpurun_out/r2u_sweep_$w.err || tail -5 gpurun_out/r2u_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2u_sweep_$w.json
done
NJODE_SEG_TPN=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "segment_thread_per_neuron_kernels_train_mode and 200" 2>&1 | grep -v "^=========     \|^  " | tail -3
w=bs_demo_200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nj_segtpn -c 2 -o gpurun_out/r2u_ncu_$w -f python bench.py --steps 1 --warmup 0 --workload $w --no-cpu-baseline --no-targets > /dev/null 2> gpurun_out/r2u_ncu_$w.err
python scripts/ncu_summary.py gpurun_out/r2u_ncu_$w.ncu-rep > gpurun_out/r2u_ncu_$w.txt 2>&1
