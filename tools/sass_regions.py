"""Splits the SASS of a kernel object into barrier-delimited segments and prints the opcode mix of each
(spills = STL/LDL): python tools/sass_regions.py obj.o [min_len]"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ins = []
for l in out.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
seg = []
start = 0
for i, (a, t) in enumerate(ins):
    if t.startswith("BAR.") or t.startswith("EXIT") or t.startswith("USETMAXREG"):
        seg.append((start, i, t.split()[0] + " " + " ".join(t.split()[1:3])))
        start = i + 1
for s, e, why in seg:
    if e - s < minlen:
        continue
    c = collections.Counter(t.split()[0].split(".")[0] for _, t in ins[s:e])
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    print(f"{ins[s][0]:#07x} n={e - s:5d} fp64={fp:4d} LDS={c['LDS']:3d} STS={c['STS']:3d} LDL={c['LDL']:3d} STL={c['STL']:3d} "
          f"MUFU={c['MUFU']:2d} CALL={c['CALL']:2d} BRA={c['BRA']:3d}  ends:{why}")
