#!/bin/bash
# round 2, run aq: LAST build -- full GPU suite, smoke, eval at 100 / 500 / 4000 paths, default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for B in 100 500 4000; do echo "== eval B=$B"; timeout 300 python scripts/eval_bench.py $B 2>/dev/null | grep "outputs stay\|evaluate"; done
timeout 900 python bench.py > gpurun_out/r2aq_bench_default.json 2> gpurun_out/r2aq_bench_default.err; wc -l gpurun_out/r2aq_bench_default.json; python scripts/bench_line.py gpurun_out/r2aq_bench_default.json
