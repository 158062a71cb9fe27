#!/bin/bash
# round 2, run ai: last build of the round -- full GPU suite, smoke, default bench (both arms), the whole-path sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2ai_reference_arm.json 2>/dev/null; wc -l gpurun_out/r2ai_reference_arm.json
timeout 900 python bench.py > gpurun_out/r2ai_bench_default.json 2> gpurun_out/r2ai_bench_default.err; wc -l gpurun_out/r2ai_bench_default.json; python scripts/bench_line.py gpurun_out/r2ai_bench_default.json
for w in physionet_synth_b50 physionet_synth_b300 physionet_synth_b600 physionet_synth_b2000; do
  timeout 600 python bench.py --steps 5 --warmup 2 --workload $w --no-cpu-baseline --no-targets > gpurun_out/r2ai_sweep_$w.json 2> gpurun_out/r2ai_sweep_$w.err || tail -5 gpurun_out/r2ai_sweep_$w.err
  python scripts/bench_line.py gpurun_out/r2ai_sweep_$w.json
done
