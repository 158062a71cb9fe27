#!/usr/bin/env python
"""per-source-line instruction / stall-sample totals of one kernel: joins `ncu --page source --csv` (SASS rows,
executed counts) with `nvdisasm -g` line info of the same cubin by instruction offset.
usage: ncu_lines.py <ncu source csv> <cubin> <mangled kernel name substring> [top]"""
import collections, csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn, inl = {}, None, False, None
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iA, iN, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
addrs = [int(r[iA], 16) for r in rows[2:] if len(r) > iN]
base = min(addrs)
cnt, smp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iN:
        continue
    off = int(r[iA], 16) - base
    key = line_of.get(off, ("?", 0))
    cnt[key] += int(r[iN] or 0); smp[key] += int(r[iS] or 0)
tot, ts = sum(cnt.values()), sum(smp.values())
srcs = {}
def text(f, l):
    if f not in srcs:
        try: srcs[f] = open("/root/repo/njode_b200/csrc/" + f).read().splitlines()
        except Exception: srcs[f] = []
    return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
print("total warp instr %d, samples %d" % (tot, ts))
order = smp.most_common(top) if len(sys.argv) > 5 and sys.argv[5] == "stall" else cnt.most_common(top)
for key, _n in order:
    n = cnt[key]
    print("%5.2f%% instr %5.2f%% stall  %s:%d  %s" % (100.0 * n / tot, 100.0 * smp[key] / max(ts, 1), key[0], key[1], text(*key)))
