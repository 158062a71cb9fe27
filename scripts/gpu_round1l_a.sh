#!/bin/bash
# r1l (a): GPU parity suite after the ABI v4 / GRU change + ncu --set full captures of the generic kernels
# (2x100 nets, masked PhysioNet-shaped model) and of the segment kernels (traffic figures for bench.py)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in bs_2x100_5k physionet_synth_b2000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nj_(fwd|bwd)_kernel' -s 2 -c 2 -o gpurun_out/prof_r1l_$w -f \
    python bench.py --steps 1 --warmup 1 --workload $w --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1; tail -2 gpurun_out/ncu_$w.log | cut -c1-200
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nj_seg_(fwd|bwd)_kernel' -s 2 -c 2 -o gpurun_out/prof_r1l_seg -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seg.log 2>&1; tail -2 gpurun_out/ncu_seg.log | cut -c1-200
ls -la gpurun_out | tail -8
