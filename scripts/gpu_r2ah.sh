#!/bin/bash
# round 2, run ah: thread-per-neuron forward stages only the jump networks' image (two CTAs per SM); class-A backward threshold
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for w in physionet_synth_b50 physionet_synth_b300 physionet_synth_b600 physionet_synth_b2000 bs_demo_gru_500 bs_demo_gru_5k heston_demo_20k bs_demo_200; do
  timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-targets > gpurun_out/r2ah_sweep_$w.json 2> gpurun_out/r2ah_sweep_$w.err || tail -5 gpurun_out/r2ah_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2ah_sweep_$w.json
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -q -x -k "every_kernel_family and tpn or thread_per_neuron" 2>&1 | grep -v "^=========     \|^  " | tail -3
