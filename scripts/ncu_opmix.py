#!/usr/bin/env python
"""opcode mix of a kernel from `ncu --page source --csv` (SASS view): executed warp instructions per mnemonic"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iN: continue
    src = r[iS].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "MUFU")) and "." in op else "")
    n = int(r[iN] or 0); mix[op] += n; tot += n; samp[op] += int(r[iSamp] or 0)
ts = sum(samp.values())
print("total warp instr", tot)
for op, n in mix.most_common(28):
    print("%-14s %14d %6.2f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))
