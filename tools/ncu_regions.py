"""Per code-region view of an `ncu --page source --csv` export: contiguous runs of SASS with the same execution
count (= one loop body of one role) with their stall samples.  python tools/ncu_regions.py source.csv [min_len]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, data = rows[1], rows[2:]
ia, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
sb = hdr.index('stall_barrier')
names = hdr[sb:sb + 17]
regs, cur = [], None
for idx, x in enumerate(data):
    e = int(x[ie])
    if cur is None or e != cur[0]:
        cur = [e, idx, idx, 0, collections.Counter(), collections.Counter()]
        regs.append(cur)
    cur[2] = idx
    cur[3] += int(x[ia])
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[1])
    cur[4][m.group(2).split('.')[0]] += 1
    for j, n in enumerate(names):
        cur[5][n[6:]] += int(x[sb + j] or 0)
tot = sum(r[3] for r in regs)
print(f"total samples {tot}")
for r in regs:
    n = r[2] - r[1] + 1
    if (n >= minlen and r[0] > 0) or r[3] > tot * 0.01:
        c = r[4]
        print(f"[{r[1]:5d}-{r[2]:5d}] n={n:4d} exec={r[0]:7d} samples={r[3]:6d} fp64={c['DFMA'] + c['DMUL'] + c['DADD']:4d} "
              f"LDS={c['LDS']:3d} STS={c['STS']:3d} LDL={c['LDL']:3d} STL={c['STL']:3d} | " +
              " ".join(f"{k}={v}" for k, v in r[5].most_common(5) if v))
