#!/bin/bash
# A/B of the segment backward's dW helper warps (NJODE_SEG_HELPERS=0/1 overrides the planner's choice)
run() { timeout 200 python bench.py --steps 10 --warmup 3 --workload $1 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['config']['workload'], 'ms', round(d['ms_per_step'],3), 'fwd', round(d['roofline'].get('fwd_kernel_ms',0),3), 'bwd', round(d['roofline']['kernel_ms'],3), d['roofline']['kernel'])"; }
for w in bs_demo_200 hestonwof_demo_1k bs_2x100_5k; do NJODE_SEG_HELPERS=0 run $w helpers=0; NJODE_SEG_HELPERS=1 run $w helpers=1; run $w planner; done
run heston_demo_20k planner
