#!/bin/bash
for r in 1 2 4 8; do NJODE_PATH_R=$r python scripts/dbg_path.py 12 40 2>&1 | grep -v "^\[dw\]\|^\[bwd\]" | tail -2; done
NJODE_PATH_R=1 python scripts/dbg_path.py 6 20 2>&1 | grep -v "^\[dw\]\|^\[bwd\]" | tail -1
NJODE_PATH_R=1 python scripts/dbg_path.py 6 40 2>&1 | grep -v "^\[dw\]\|^\[bwd\]" | tail -1
NJODE_PATH_R=1 python scripts/dbg_path.py 12 20 2>&1 | grep -v "^\[dw\]\|^\[bwd\]" | tail -1
