"""Builds A/B variants of the library next to the product build:
    python tools/build_variants.py name:DEF1,DEF2 name2:DEF3 ...
-> dvi_ekf_b200/libeskf_b200_<name>.so with -D<DEF> each (select at run time with ESKF_B200_LIB)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvi_ekf_b200 import build  # noqa: E402

for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    here, obj, lib = build.HERE, build.OBJ, build.LIB
    print(build.build_cuda(defines=tuple(d for d in defs.split(",") if d), suffix="_" + name), flush=True)
    log = os.path.join(build.OBJ, "eskf_launch3_f28.o.log")
    for l in open(log).read().splitlines():
        if "spill" in l and "bytes stack" in l:
            print("   ", l.strip())
            break
    build.OBJ, build.LIB = obj, lib
